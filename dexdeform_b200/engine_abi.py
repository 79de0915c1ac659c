"""argtypes of the fused engine entry points (ABI-2, ``dd_*`` in include/dexdeform_mpm.h)."""


def bind_abi2(library):
    return library
