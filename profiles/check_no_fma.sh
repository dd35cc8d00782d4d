#!/usr/bin/env bash
# Build-time parity guard: the exact-reciprocal (RECIP=true) kernels must not contain a single fused
# multiply-add -- contraction changes results, and ptxas 12.9 was seen contracting packed
# mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false. The IEEE-division variants (RECIP=false)
# legitimately contain FFMA inside the division sequence and are not checked.
set -euo pipefail
LIB="${1:-$(dirname "$0")/../nvidia-jetson-workload_b200/lib/libweather_b200.so}"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3; recip = (fn ~ /step_(tma|fused)_kernelI[fd]Li[0-9]+ELi[0-9]+ELb1/) || (fn ~ /(stage_direct|diagnostics)_kernelI[fd]Lb1/); if (recip) checked++ }
/FFMA|DFMA/ { if (recip) { bad[fn]++ } }
END { n=0; for (f in bad) { print "FMA in exact-reciprocal kernel: " bad[f] " x " f; n++ }
      if (checked == 0) { print "check_no_fma: no kernels matched"; exit 2 }
      if (n) exit 1; print "check_no_fma: ok (" checked " exact-reciprocal kernels, 0 fused multiply-adds)" }'
