#!/bin/bash
# Per-kernel SASS mnemonic counts of the shipped library (evidence for which hardware paths the kernels use):
#   DMMA.8x8x4 (FP64 tensor pipe), UTMALDG (TMA tensor loads), SYNCS (mbarrier), LDGSTS (cp.async), MUFU (SFU seeds)
# usage: tools/sass_summary.sh > profiles/r02/sass_summary.txt
so=${1:-bobe_b200/lib/libbobe_b200.so}
echo "# cuobjdump -sass $so  ($(date -u +%F)), nvcc $(/usr/local/cuda/bin/nvcc --version | grep release | sed 's/.*release //')"
/usr/local/cuda/bin/cuobjdump -sass "$so" | awk '
/Function :/ { name=$3; order[++n]=name }
/DMMA/ { dmma[name]++ }
/UTMALDG/ { tma[name]++ }
/SYNCS/ { syncs[name]++ }
/LDGSTS/ { ldgsts[name]++ }
/MUFU/ { mufu[name]++ }
/DFMA|DADD|DMUL/ { dfma[name]++ }
/UTCHMMA|UTCMMA|tcgen05/ { utc[name]++ }
END { printf "%-8s %-8s %-7s %-7s %-6s %-7s %-5s %s\n","DMMA","UTMALDG","SYNCS","LDGSTS","MUFU","DFMA..","UTC*","kernel";
      for(i=1;i<=n;i++){k=order[i]; printf "%-8d %-8d %-7d %-7d %-6d %-7d %-5d %s\n",dmma[k],tma[k],syncs[k],ldgsts[k],mufu[k],dfma[k],utc[k],k} }' | (read h; echo "$h"; sort -k1,1nr) | c++filt | cut -c1-220
