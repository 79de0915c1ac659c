// inert stand-in: only the (out-of-scope) renderer kernel uses cuRAND
#pragma once
