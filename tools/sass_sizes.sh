#!/bin/bash
# Code size (bytes) of every out-of-line subroutine inside one instantiation of the step kernel, from nvdisasm labels.
# usage: tools/sass_sizes.sh [kernel template args, default 2ELi1ELi1E]
K=${1:-2ELi1ELi1E}
cd /tmp && rm -f ncg_b200.sm_100a.cubin && cuobjdump -xelf all /root/repo/nascargymnasium_b200/libncg_b200.so >/dev/null 2>&1
nvdisasm ncg_b200.sm_100a.cubin 2>/dev/null > /tmp/all_dis.txt
python3 - "$K" <<'PY'
import re, sys
K = sys.argv[1]
lines = open('/tmp/all_dis.txt').read().splitlines()
inside = False; cur = 'kernel body'; sizes = {}; last_addr = 0; start = {}
for ln in lines:
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m:
        inside = ('ncg_step_kernelILi' + K) in m.group(1); cur = 'kernel body'; continue
    if not inside: continue
    m = re.match(r'^\$\S+\$(_Z\S+):', ln)
    if m: cur = m.group(1); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/', ln)
    if m:
        sizes[cur] = sizes.get(cur, 0) + 16
tot = sum(sizes.values())
for k, v in sorted(sizes.items(), key=lambda kv: -kv[1]):
    print(f"{v:8d} B  {k[:100]}")
print(f"{tot:8d} B  total")
PY
