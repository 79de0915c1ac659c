#!/bin/bash
# static SASS statistics of the hot kernels of a built library: instructions, registers, local-memory traffic
LIB=${1:-dexdeform_b200/libmaniskill_mpm.so}
cuobjdump -sass $LIB 2>/dev/null | awk '/Function : /{name=$3} /\/\*[0-9a-f]{4}\*\//{n[name]++; if ($0 ~ /LDL|STL/) l[name]++; if ($0 ~ /LDS/) s[name]++; if ($0 ~ /STS/) t[name]++} END{for (k in n) if (k ~ /k_p2g_tile|k_g2p_grad_tile|k_g2p_tile|k_p2g_grad_tile|k_grid_b|k_grid_grad_b/) printf "%-90s instr %5d  local %3d  lds %3d sts %3d\n", substr(k,1,90), n[k], l[k], s[k], t[k]}' | sort
cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 -E "k_p2g_tile|k_g2p_grad_tile|k_g2p_tile|k_p2g_grad_tile|k_grid_bE|k_grid_grad_b" | grep -E "Function|REG" | paste - - | sed -E 's/Function ([^:]*):.*REG:([0-9]+) STACK:([0-9]+).*/\1 REG \2 STACK \3/' | cut -c1-140
