/* cmf_b200_tools.h -- extra entry point of the TOOLS build (libcmf_b200_tools.so) of the matched-filter library:
 * the same sources as libcmf_b200.so compiled with -DCMF_TUNING_HOOKS (environment tuning / cross-check hooks, the
 * Jacobi cross-check solver, kernel variants for the tuning sweeps) plus the micro-benchmarks that measure the
 * roofline denominators recorded under profiles/.  Nothing here is part of the product ABI (include/cmf_b200.h). */
#ifndef CMF_B200_TOOLS_H
#define CMF_B200_TOOLS_H

#include "cmf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- micro-benchmarks used for the roofline denominators (profiles/): returns achieved rate ---- */
/* kind: 0 DMMA.8x8x4 TFLOP/s (32 warps/SM), 8 same with 8 warps/SM and 24 accumulators, 1 DFMA TFLOP/s,
 *       2 HBM read GB/s (8-byte loads), 3 HBM read GB/s (16-byte), 4 HBM copy GB/s (read+write bytes),
 *       5 bulk-async-copy read GB/s, 6 f32->f64 conversions G/s, 7 FP64 log+divide pairs G/s,
 *       9 legacy mma.sync tf32 TFLOP/s, 10 legacy mma.sync bf16 TFLOP/s, 11 FP32 FFMA TFLOP/s,
 *       12-15 dependent-issue latency in cycles of DFMA / rsqrt+DADD / sqrt+DADD / divide+DADD (one thread),
 *       16-18 cycles per step of the QL rotation recurrence: alone / 20 chains per SM / 5 chains per SM,
 *       40-56 cycles per TS-form tcgen05 TF32 MMA (M = 128, K = 8) with N = 16 (kind - 40) columns (40: N = 256),
 *             27-MMA rounds into one accumulator issued back to back by one CTA,
 *       60-63 cycles per warp instruction and scheduler of FFMA (60: 1 warp per scheduler, 61: 4) and of the packed
 *             FFMA2 (62, 63), 16 independent chains per thread,
 *       30-33 GB/s of the two halves of the repack pass alone: slab read through 4-byte LDGSTS / 8-byte LDGSTS /
 *             8-byte loads, and the 16-byte xt write side */
double cmf_microbench(int device, int kind, int iters);

/* SM clocks at the hand-offs (convert, GEMM1, square, GEMM2, epilogue) of four consecutive 128-pixel tiles of one CTA
 * of the last loo_screen5_kernel launch: out[4][32], 0 where an event was not reached (tools/s5_timeline.py). */
int cmf_tools_s5_timeline(long long* out);

#ifdef __cplusplus
}
#endif
#endif /* CMF_B200_TOOLS_H */
