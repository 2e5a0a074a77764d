// Micro-benchmarks behind cmf_microbench(): the measured denominators used in DESIGN.md / profiles/
// for the FP64 tensor path (MEASURED_PEAKS.json only carries HBM copy and bf16 GEMM figures).
#include <stdio.h>

#include "../../include/cmf_b200_tools.h"
#include "cmf_common.cuh"
#include "cmf_internal.h"

namespace {

using namespace cmf;

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters) {
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = i;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) cvt_kernel(double* out, const float* in, int iters) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = in[threadIdx.x + i];
    double s = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s += (double)x[i];
            x[i] = __int_as_float(__float_as_int(x[i]) + 1);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) logdiv_kernel(double* out, int iters) {
    double r = 50.0 + threadIdx.x * 0.01, s = 0.0;
    const double beta = 5e-5;
    for (int it = 0; it < iters; ++it) {
        const double q = 1.0 - beta * r;
        s += log(q) + r / q;
        r += 1e-3;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// legacy warp-level tensor path (mma.sync, SASS HMMA): how much of it survives on sm_100a
template <int NACC>
__global__ void __launch_bounds__(256) mma_tf32_kernel(float* out, int iters) {
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
    uint32_t a0 = 0x3f800000u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) mma_bf16_kernel(float* out, int iters) {
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
    uint32_t a0 = 0x3f803f80u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) ffma_kernel(float* out, int iters) {
    float c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    const float a = 1.0f + threadIdx.x * 1e-7f, b = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename V>
__global__ void __launch_bounds__(256) read_kernel(const V* __restrict__ in, size_t n, float* out) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        V v0 = in[i], v1 = in[i + stride], v2 = in[i + 2 * stride], v3 = in[i + 3 * stride];
        const float* f0 = reinterpret_cast<const float*>(&v0);
        const float* f1 = reinterpret_cast<const float*>(&v1);
        const float* f2 = reinterpret_cast<const float*>(&v2);
        const float* f3 = reinterpret_cast<const float*>(&v3);
        acc += f0[0] + f1[0] + f2[0] + f3[0];
    }
    for (; i < n; i += stride) { V v = in[i]; acc += reinterpret_cast<const float*>(&v)[0]; }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void __launch_bounds__(256) copy_kernel(const float4* __restrict__ in, float4* __restrict__ outp, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) outp[i] = in[i];
}

// every warp streams 8 KB tiles through a 3-stage ring with 1-D bulk async copies
__global__ void __launch_bounds__(256) bulk_read_kernel(const float* __restrict__ in, size_t ntiles, float* out) {
    constexpr int TB = 2048, NS = 3;   // floats per tile (8 KB)
    extern __shared__ __align__(128) unsigned char sraw[];
    float* ring = reinterpret_cast<float*>(sraw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 8 * NS * TB);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { for (int s = 0; s < NS; ++s) mbar_init(&bars[warp * NS + s], 1); fence_mbar_init(); }
    __syncthreads();
    const size_t gw = (size_t)blockIdx.x * 8 + warp, nw = (size_t)gridDim.x * 8;
    float* my = ring + warp * NS * TB;
    auto issue = [&](size_t it) {
        const size_t t = gw + nw * it;
        if (t < ntiles) {
            mbar_expect_tx(&bars[warp * NS + it % NS], TB * 4);
            bulk_g2s(my + (it % NS) * TB, in + t * TB, TB * 4, &bars[warp * NS + it % NS]);
        }
    };
    if (lane == 0) for (int it = 0; it < NS; ++it) issue(it);
    float acc = 0.f;
    for (size_t it = 0; gw + nw * it < ntiles; ++it) {
        mbar_wait(&bars[warp * NS + it % NS], (uint32_t)((it / NS) & 1));
        acc += my[(it % NS) * TB + lane * 4];
        __syncwarp();
        if (lane == 0) issue(it + NS);
    }
    if (acc == 123.456f) out[0] = acc;
}

float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

}  // namespace

// dependent-issue latency in cycles of one FP64 op, one thread of one warp (kinds 12-15): the serial rotation
// recurrence of the QL eigen-solver is a chain of such ops, so its length is what bounds K2
template <int OP>
__global__ void fp64_latency_kernel(double* out, long long* cycles, int n, double seed) {
    double x = seed, y = 1.0 + seed * 1e-3;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) x = fma(x, y, 1e-9);            // DFMA
            else if (OP == 1) x = rsqrt(x) + 1.5;        // rsqrt + DADD
            else if (OP == 2) x = sqrt(x) + 2.0;         // sqrt + DADD
            else x = 1.0 / x + 0.5;                      // divide + DADD
        }
    }
    const long long t1 = clock64();
    out[0] = x;
    cycles[0] = t1 - t0;
}

// the rotation recurrence of the QL eigen-solver (k_eigen.cu, tql2 producer) in isolation: cycles per step of one
// lane, with `blockDim.x / 32` warps per CTA all running their own copy (kind 16: one warp on one SM; kind 17: 20 warps
// per SM like 5 resident eigen CTAs, every warp a chain; kind 18: 5 chains per SM, one per CTA)
__global__ void ql_chain_kernel(double* out, long long* cycles, int n, int steps) {
    __shared__ double dsh[8][80], esh[8][80], csh[8][160];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* d = dsh[w]; double* e = esh[w]; double* cs = csh[w];
    if (lane == 0) for (int i = 0; i < 80; ++i) { d[i] = 1.0 + 0.01 * i; e[i] = 0.3 + 0.001 * i; }
    __syncwarp();
    long long t0 = 0, t1 = 0;
    if (lane == 0) {
        t0 = clock64();
        double g = 0.37, sn = 1.0, c = 1.0, p = 0.0;
        for (int rep = 0; rep < n; ++rep) {
            int i = steps - 1, cnt = 0;
            double e_i = e[i], d_i = d[i], d_i1 = d[i + 1];
            bool under = false;
            for (; i >= 0 && !under; --i) {
                double e_n = 0.0, d_n = 0.0;
                if (i > 0) { e_n = e[i - 1]; d_n = d[i - 1]; }
                const double f = sn * e_i, b = c * e_i;
                const double h = f * f + g * g;
                under = !(h > 0.0);
                double x0;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(h));
                const double err = fma(-h * x0, x0, 1.0);
                const double rinv = under ? 0.0 : fma(fma(err, 0.375, 0.5), x0 * err, x0);
                const double g2 = d_i1 - p;
                const double wv = fma(d_i - g2, f, 2.0 * g * b);
                e[i + 1] = h * rinv;
                sn = f * rinv; c = g * rinv;
                const double r2 = wv * rinv;
                p = sn * r2;
                d[i + 1] = g2 + p;
                g = c * r2 - b;
                cs[2 * cnt] = c; cs[2 * cnt + 1] = sn; ++cnt;
                e_i = e_n; d_i1 = d_i; d_i = d_n;
            }
            g = 0.37 + 1e-3 * g; sn = 1.0; c = 1.0; p = 0.0;
        }
        t1 = clock64();
        out[blockIdx.x * 8 + w] = g + d[3];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}

// The two halves of the repack pass in isolation (kinds 30-33), same shapes as repack_pipe_kernel<9, 32, 4, 4>:
//   READ : slab [L][72][598] -> shared, runs of 32 columns (128 B) per (line, band) row, 4-line tiles, ring of 4,
//          through 4-byte LDGSTS (kind 30), 8-byte LDGSTS (31) or plain 8-byte loads to registers (32)
//   WRITE: 16-byte stores of each column's contiguous [4][72] block of xt [598][L][72] (33)
template <int MODE>
__global__ void __launch_bounds__(576) repack_half_kernel(const float* __restrict__ slab, float* __restrict__ xt,
                                                          int L, int S, int D, int lines_per_cta, float* sink) {
    constexpr int CG = 32, LT = 4, NS = 4, DP = 72, CS = LT * DP + 4;
    extern __shared__ __align__(16) float tile[];
    const int tid = threadIdx.x, s0 = blockIdx.x * CG;
    const int l_begin = blockIdx.y * lines_per_cta, l_end = min(L, l_begin + lines_per_cta);
    const int ntiles = (l_end - l_begin + LT - 1) / LT;
    float accf = 0.f;
    if (MODE <= 2) {
        const int w = (MODE == 0) ? 1 : 2;                       // columns per copy
        const int c_ld = (tid % (CG / w)) * w, r_ld = tid / (CG / w);
        const int rows_per_pass = 576 / (CG / w);
        const bool ok = s0 + c_ld < S;
        for (int it = 0; it < ntiles + NS - 1; ++it) {
            if (it < ntiles && ok) {
                const int l0 = l_begin + it * LT, buf = it % NS;
                for (int r = r_ld; r < LT * D; r += rows_per_pass) {
                    const int l = r / D, b = r - l * D;
                    if (l0 + l >= l_end) break;
                    const float* src = slab + ((long long)(l0 + l) * D + b) * S + s0 + c_ld;
                    float* dst = tile + (size_t)buf * CG * CS + (size_t)(r * CG + c_ld);
                    if (MODE == 0)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
                    else if (MODE == 1)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
                    else {
                        const float2 v = *reinterpret_cast<const float2*>(src);
                        accf += v.x + v.y;
                    }
                }
            }
            if (MODE != 2) {
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 3;" ::: "memory");
                __syncthreads();
            }
        }
        if (MODE != 2) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            accf = tile[tid];
        }
    } else {
        const int c = tid / 18, q = tid % 18;
        if (s0 + c < S) {
            for (int it = 0; it < ntiles; ++it) {
                const int l0 = l_begin + it * LT;
                float* dst = xt + ((long long)(s0 + c) * L + l0) * DP + 4 * q;
#pragma unroll
                for (int l = 0; l < LT; ++l)
                    if (l0 + l < l_end) *reinterpret_cast<float4*>(dst + l * DP) = make_float4(1.f, 2.f, 3.f, (float)it);
            }
        }
    }
    if (accf == 123.456f) sink[0] = accf;
}

extern "C" double cmf_microbench(int device, int kind, int iters) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double result = -1.0;
    if (iters <= 0) iters = 1;
    if (kind == 0 || kind == 1 || kind == 6 || kind == 7 || kind == 8) {
        const int blocks = sms * ((kind == 8) ? 1 : 4);
        double* out; cudaMalloc(&out, (size_t)blocks * 256 * sizeof(double));
        float* fin; cudaMalloc(&fin, 4096); cudaMemset(fin, 0, 4096);
        const int inner = 4096;
        auto run = [&]() {
            if (kind == 0) dmma_kernel<8><<<blocks, 256>>>(out, inner);
            else if (kind == 8) dmma_kernel<24><<<blocks, 256>>>(out, inner);
            else if (kind == 1) dfma_kernel<<<blocks, 256>>>(out, inner);
            else if (kind == 6) cvt_kernel<<<blocks, 256>>>(out, fin, inner);
            else logdiv_kernel<<<blocks, 256>>>(out, inner);
        };
        run(); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) run();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        const double sec = time_ms(e0, e1) * 1e-3 / iters;
        const double threads = (double)blocks * 256;
        if (kind == 0) result = threads / 32 * inner * 8 * 512.0 / sec * 1e-12;         // TFLOP/s
        else if (kind == 8) result = threads / 32 * inner * 24 * 512.0 / sec * 1e-12;   // TFLOP/s, 8 warps/SM
        else if (kind == 1) result = threads * inner * 8 * 2.0 / sec * 1e-12;           // TFLOP/s
        else if (kind == 6) result = threads * inner * 8 / sec * 1e-9;                   // Gcvt/s
        else result = threads * inner / sec * 1e-9;                                      // G (log+div)/s
        cudaFree(out); cudaFree(fin);
    } else if (kind >= 9 && kind <= 11) {
        const int blocks = sms * 4;
        float* out; cudaMalloc(&out, (size_t)blocks * 256 * sizeof(float));
        const int inner = 4096;
        auto run = [&]() {
            if (kind == 9) mma_tf32_kernel<8><<<blocks, 256>>>(out, inner);
            else if (kind == 10) mma_bf16_kernel<8><<<blocks, 256>>>(out, inner);
            else ffma_kernel<<<blocks, 256>>>(out, inner);
        };
        run(); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) run();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        const double sec = time_ms(e0, e1) * 1e-3 / iters;
        const double threads = (double)blocks * 256;
        if (kind == 9) result = threads / 32 * inner * 8 * (2.0 * 16 * 8 * 8) / sec * 1e-12;        // TFLOP/s
        else if (kind == 10) result = threads / 32 * inner * 8 * (2.0 * 16 * 8 * 16) / sec * 1e-12;  // TFLOP/s
        else result = threads * inner * 16 * 2.0 / sec * 1e-12;                                      // TFLOP/s
        cudaFree(out);
    } else if (kind >= 2 && kind <= 5) {
        const size_t bytes = (size_t)4 << 30;   // 4 GiB, far larger than the 126 MB L2
        float* in; float* outp = nullptr; float* flag;
        if (cudaMalloc(&in, bytes) != cudaSuccess) return -1.0;
        cudaMalloc(&flag, 256);
        cudaMemset(in, 0, bytes);
        if (kind == 4) { if (cudaMalloc(&outp, bytes) != cudaSuccess) { cudaFree(in); return -1.0; } }
        const int blocks = sms * 8;
        if (kind == 5) cudaFuncSetAttribute(bulk_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 8192 + 256);
        auto run = [&]() {
            if (kind == 2) read_kernel<float2><<<blocks, 256>>>(reinterpret_cast<const float2*>(in), bytes / 8, flag);
            else if (kind == 3) read_kernel<float4><<<blocks, 256>>>(reinterpret_cast<const float4*>(in), bytes / 16, flag);
            else if (kind == 4) copy_kernel<<<blocks, 256>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(outp), bytes / 16);
            else bulk_read_kernel<<<sms, 256, 8 * 3 * 8192 + 256>>>(in, bytes / 8192, flag);
        };
        run(); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) run();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        const double sec = time_ms(e0, e1) * 1e-3 / iters;
        result = (kind == 4 ? 2.0 : 1.0) * (double)bytes / sec * 1e-9;   // GB/s
        cudaFree(in); cudaFree(flag); if (outp) cudaFree(outp);
    }
    else if (kind >= 12 && kind <= 15) {
        double* o; long long* c;
        cudaMalloc(&o, 8); cudaMalloc(&c, 8);
        const int n = 4096;
        for (int rep = 0; rep < 2; ++rep) {
            if (kind == 12) fp64_latency_kernel<0><<<1, 1>>>(o, c, n, 1.0000001);
            else if (kind == 13) fp64_latency_kernel<1><<<1, 1>>>(o, c, n, 1.7);
            else if (kind == 14) fp64_latency_kernel<2><<<1, 1>>>(o, c, n, 1.7);
            else fp64_latency_kernel<3><<<1, 1>>>(o, c, n, 1.7);
        }
        long long cyc = 0;
        cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
        result = (double)cyc / (8.0 * n);                  // cycles per op (kinds 13-15 include one DADD)
        cudaFree(o); cudaFree(c);
    }
    else if (kind >= 16 && kind <= 18) {
        double* o; long long* c;
        cudaMalloc(&o, 8 * 8 * 148 * 8); cudaMalloc(&c, 8);
        const int n = 64, steps = 70;
        for (int rep = 0; rep < 2; ++rep) {
            if (kind == 16) ql_chain_kernel<<<1, 32>>>(o, c, n, steps);
            else if (kind == 17) ql_chain_kernel<<<sms * 5, 128>>>(o, c, n, steps);
            else ql_chain_kernel<<<sms * 5, 32>>>(o, c, n, steps);
        }
        long long cyc = 0;
        cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
        result = (double)cyc / ((double)n * steps);        // cycles per rotation step
        cudaFree(o); cudaFree(c);
    }
    else if (kind >= 30 && kind <= 33) {
        const int L = 20000, S = 598, D = 72;
        const size_t n = (size_t)L * S * D;
        float *slab, *xt, *sink;
        if (cudaMalloc(&slab, n * 4) != cudaSuccess) return -1.0;
        if (cudaMalloc(&xt, n * 4) != cudaSuccess) { cudaFree(slab); return -1.0; }
        cudaMalloc(&sink, 16);
        cudaMemset(slab, 0, n * 4);
        const size_t smem = (size_t)4 * 32 * (4 * 72 + 4) * 4;
        const dim3 grid(19, 7);
        const int lpc = (L + 6) / 7;
        auto run = [&]() {
            if (kind == 30) repack_half_kernel<0><<<grid, 576, smem>>>(slab, xt, L, S, D, lpc, sink);
            else if (kind == 31) repack_half_kernel<1><<<grid, 576, smem>>>(slab, xt, L, S, D, lpc, sink);
            else if (kind == 32) repack_half_kernel<2><<<grid, 576, smem>>>(slab, xt, L, S, D, lpc, sink);
            else repack_half_kernel<3><<<grid, 576, smem>>>(slab, xt, L, S, D, lpc, sink);
        };
        cudaFuncSetAttribute(repack_half_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(repack_half_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(repack_half_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(repack_half_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        run(); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) run();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        const double sec = time_ms(e0, e1) * 1e-3 / iters;
        result = (double)n * 4 / sec * 1e-9;                  // GB/s of the 3.44 GB slab (read) or xt (write)
        cudaFree(slab); cudaFree(xt); cudaFree(sink);
    }
    else if (kind >= 20 && kind <= 25) {
        // tcgen05 TS-form self test: max |D - A.B^T| (exact inputs, so 0 when the layouts are right)
        if (kind == 20) result = cmf::screen5_selftest(80, 72, 0, 0);
        else if (kind == 21) result = cmf::screen5_selftest(80, 72, 0, 1);
        else if (kind == 22) result = cmf::screen5_selftest(96, 72, 112, 0);
        else if (kind == 23) result = cmf::screen5_selftest(112, 72, 0, 0);
        else if (kind == 24) result = cmf::screen5_selftest(72, 72, 0, 0);
        else result = cmf::screen5_selftest(256, 8, 0, 0);
    }
    else if (kind >= 40 && kind <= 56) {
        // cycles per TS-form TF32 MMA with N = 16 (kind - 40) columns (kind 40: N = 256)
        result = cmf::screen5_mma_rate(kind == 40 ? 256 : 16 * (kind - 40), iters);
    }
    else if (kind >= 60 && kind <= 63) {
        // cycles per FFMA (60, 61) / FFMA2 (62, 63) warp instruction and scheduler, 1 or 4 warps per scheduler
        result = cmf::fma_issue_rate(kind >= 62, (kind & 1) ? 4 : 1);
    }
    if (cudaGetLastError() != cudaSuccess) result = -1.0;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return result;
}

// SM clocks at the hand-offs of four consecutive 128-pixel tiles of one CTA of the last loo_screen5_kernel launch
// (tools/s5_timeline.py): out[4][32], 0 where an event was not reached
extern "C" int cmf_tools_s5_timeline(long long* out) { return out ? cmf::screen5_timeline(out) : -1; }
