#!/usr/bin/env bash
# Build-time parity guard: the exact-reciprocal (RECIP=true) kernels must not contain a single CONTRACTED
# multiply-add -- contraction changes results, and ptxas 12.9 was seen contracting packed
# mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false.
# Allowed, because exact by construction (wsb_arith.cuh): FFMA2 whose multiplier is the IMMEDIATE -1 (the packed
# subtraction c - p) or, in the folded-arithmetic instantiations only (last template flag), the immediate 2.
# Any FFMA/FFMA2/DFMA with a register (or any other constant) multiplier fails the build.
# The IEEE-division variants (RECIP=false) legitimately contain FFMA inside the division sequence: not checked.
set -euo pipefail
LIB="${1:-$(dirname "$0")/../nvidia-jetson-workload_b200/lib/libweather_b200.so}"
cuobjdump -sass "$LIB" | awk '
/Function :/ {
    fn = $3
    recip = (fn ~ /step_(tma|fused)_kernelI[fd]Li[0-9]+ELi[0-9]+ELb1/) || (fn ~ /(stage_direct|diagnostics|tracer_stage)_kernelI[fd]Lb1/)
    folded = (fn ~ /step_tma_kernelIfLi[0-9]+ELi[0-9]+ELb1ELi[0-9]+ELb0ELb1E/)
    if (recip) checked++
}
/FFMA|DFMA/ {
    if (!recip) next
    if ($0 ~ /FFMA2 [^;]*, -1, /) { exact++; next }
    if (folded && $0 ~ /FFMA2 [^;]*, 2, /) { exact++; next }
    bad[fn]++
}
END { n=0; for (f in bad) { print "contracted FMA in exact-reciprocal kernel: " bad[f] " x " f; n++ }
      if (checked == 0) { print "check_no_fma: no kernels matched"; exit 2 }
      if (n) exit 1
      print "check_no_fma: ok (" checked " exact-reciprocal kernels, 0 contracted multiply-adds, " exact+0 " exact FFMA2 with immediate -1 / 2)" }'
