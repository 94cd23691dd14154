#!/bin/bash
# Dev tool (here): static SASS size per device function of one kernel of the built library
# usage: tools/sass_sizes.sh <kernel name substring> [library]
K=${1:-rc_phys_kernel}
LIB=${2:-/root/repo/dynenv_b200/libdynenv_b200.so}
rm -rf /tmp/cub_sizes && mkdir -p /tmp/cub_sizes && cd /tmp/cub_sizes && cuobjdump -xelf all "$LIB" > /dev/null 2>&1
(for f in *.cubin; do nvdisasm -c $f; done) 2>/dev/null | awk -v K="$K" '
/^\/\/-+ \.text\./ {insec = (index($0, K) > 0); fn="(kernel body)"; next}
/^\/\/-+/ {insec=0}
insec && /^\$[_A-Za-z0-9$]+:/ {fn=$1}
insec && /^ +\/\*[0-9a-f]+\*\// {c[fn]++}
END {t=0; for (f in c) {print c[f]*16/1024 " KB", f; t+=c[f]} print t*16/1024 " KB total"}' | sort -n | sed 's/\$_ZN46_GLOBAL__N__[0-9a-z_]*dynenv_api_cu_[0-9a-f]*[0-9A-Za-z_]*\$_ZN6dynenv//' | cut -c1-100
