// Issue-rate microbenchmarks that give the SAD kernel its roofline denominator (SURVEY 8d:
// "measure, don't assume" -- the int32 lanes per SM of sm_100 are not documented here).
//   mode 0: independent VABSDIFF (|a-b|+c) chains                 -> peak integer SAD px-ops/s
//   mode 1: FADD pairs (d = a-b; acc += |d|)                      -> the same op on the fp32 pipe
//   mode 2: 8 VABSDIFF + 4 FADD pairs per step                    -> both pipes co-issued
//   mode 3: 8 VABSDIFF + 8 FADD pairs per step
//   modes 4..9: 8 VABSDIFF + 2 x {LDS.32, IMAD reg, IMAD imm, FADD, IADD3, VIMNMX3} per step -- what does an extra
//               instruction cost next to the SAD stream? (px_ops counts the 8 SADs only)
#include <stdlib.h>

#include "memci_internal.cuh"

template <int MODE>
__global__ void __launch_bounds__(512) ub_kernel(int iters, const int *__restrict__ seed, unsigned *__restrict__ out)
{
    int r[8];
    unsigned a[8];
    float fr[8], fa[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        r[j] = seed[(threadIdx.x + j) & 63] + j;
        a[j] = (unsigned)seed[(threadIdx.x + 8 + j) & 63];
        fr[j] = (float)r[j];
        fa[j] = (float)a[j];
    }
    __shared__ int ub_sm[2048];
    unsigned extra = 0, e0 = threadIdx.x, e1 = threadIdx.x * 3u, e2 = 7u, e3 = 11u;
    float f0 = (float)threadIdx.x, f1 = 1.5f;
    if (MODE == 4 || (MODE >= 10 && MODE != 15)) { for (int i = threadIdx.x; i < 2048; i += blockDim.x) ub_sm[i] = seed[i & 63]; __syncthreads(); }
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        if (MODE == 4 || MODE == 16) {          // 2 LDS.32 with immediate offsets (no address arithmetic), results consumed by the SAD stream
            const int *bp = ub_sm + threadIdx.x;
            r[0] = bp[(it & 3) * 32];
            r[1] = bp[512 + (it & 3) * 32];
        }
        if (MODE == 10) {         // 1 LDS.128
            const int4 v4 = *reinterpret_cast<const int4 *>(ub_sm + (threadIdx.x & 127) * 4 + (it & 3) * 512);
            r[0] = v4.x + (it & 3); r[1] = v4.y; r[2] = v4.z; r[3] = v4.w;
        }
        if (MODE == 11) {         // 4 LDS.32
            const int *bp = ub_sm + threadIdx.x;
            r[0] = bp[(it & 3) * 32]; r[1] = bp[512 + (it & 3) * 32]; r[2] = bp[1024 + (it & 3) * 32]; r[3] = bp[1536 - 512 + (it & 3) * 32 + 128];
        }
        if (MODE == 5) {
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(e0) : "r"(e2), "r"(e3));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(e1) : "r"(e3), "r"(e2));
        }
        if (MODE == 6) {
            asm volatile("mad.lo.u32 %0, %1, 0x10624dd3, %0;" : "+r"(e0) : "r"(e2));
            asm volatile("mad.lo.u32 %0, %1, 0x3e8, %0;" : "+r"(e1) : "r"(e3));
        }
        if (MODE == 7) { f0 = __fadd_rn(f0, f1); f1 = __fadd_rn(f1, f0); }
        if (MODE == 8) {
            asm volatile("add.u32 %0, %0, %1;" : "+r"(e0) : "r"(e2));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(e1) : "r"(e3));
        }
        if (MODE == 9) {
            asm volatile("min.u32 %0, %0, %1;" : "+r"(e0) : "r"(e2 + it));
            asm volatile("min.u32 %0, %0, %1;" : "+r"(e1) : "r"(e3 + it));
        }
        if (MODE == 15 || MODE == 16) {   // packed 8-bit form: VABSDIFF4.U8.ACC = 4 pixel pairs per instruction
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __vsadu4((unsigned)r[j], a[(j + 1) & 7]) + a[j];
        } else if (MODE == 0 || MODE >= 2) {   // the SAD stream (modes 4, 10, 11 refresh part of r[] from shared memory)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __sad(r[j], (int)a[(j + 1) & 7], a[j]);
        }
        if (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j) fa[j] = __fadd_rn(fa[j], fabsf(__fsub_rn(fr[j], fa[(j + 1) & 7])));
        }
        if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) fa[j] = __fadd_rn(fa[j], fabsf(__fsub_rn(fr[j], fa[(j + 1) & 3])));
        }
    }
    extra = e0 + e1 + (unsigned)f0 + (unsigned)f1;
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j] + (unsigned)fa[j];
    s += extra;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 12: the full-search kernel's own SAD pattern -- 64 reference registers, G = 11 accumulators, G + 7 window rows
// of 8 pixels per group, pixel-major order -- with the window pixels synthesised in registers (no shared memory) and
// no bookkeeping: what the register-tiled stream itself can reach at the kernel's occupancy (512 threads, 1 CTA / SM).
__global__ void __launch_bounds__(512, 1) ub_fs_pattern_kernel(int iters, const int *__restrict__ seed, unsigned *__restrict__ out)
{
    constexpr int G = 11;
    int r[8][8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) r[k][j] = seed[(threadIdx.x + k * 8 + j) & 63] + k;
    unsigned total = 0;
    int base = seed[threadIdx.x & 63];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        unsigned acc[G];
#pragma unroll
        for (int i = 0; i < G; ++i) acc[i] = 0xffffffffu;
#pragma unroll
        for (int w = 0; w < G + 7; ++w) {
            int px[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) px[j] = base + w * 8 + j + it;
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    const int kk = w - gi;
                    if (kk >= 0 && kk < 8) acc[gi] = __sad(r[kk][j], px[j], acc[gi]);
                }
        }
        unsigned m = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < G; ++i) m = min(m, acc[i]);
        total += m;
        base += (int)(m & 3u);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
}

extern "C" MEMCI_API int memci_microbench(memci_ctx *ctx, int mode, int iters, float *ms, double *px_ops)
{
    if (!ctx || mode < 0 || mode > 16 || iters <= 0) return MEMCI_ERR_ARG;
    cudaSetDevice(ctx->device);
    // modes 13 / 14: the mode-12 kernel with 2 / 1 warps per scheduler
    int grid = (mode >= 12 && mode <= 14) ? ctx->num_sms : ctx->num_sms * 4, nt = mode == 13 ? 256 : mode == 14 ? 128 : 512;
    if (const char *e = getenv("MEMCI_UB_CTAS_PER_SM")) grid = ctx->num_sms * atoi(e);        // occupancy experiments
    int *d_seed = nullptr;
    unsigned *d_out = nullptr;
    int h_seed[64];
    for (int i = 0; i < 64; ++i) h_seed[i] = (i * 7919 + 13) % 255000;
    CK(ctx, cudaMalloc(&d_seed, sizeof(h_seed)));
    CK(ctx, cudaMalloc(&d_out, sizeof(unsigned) * (size_t)grid * nt));
    CK(ctx, cudaMemcpyAsync(d_seed, h_seed, sizeof(h_seed), cudaMemcpyHostToDevice, ctx->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        if (rep == 1) cudaEventRecord(e0, ctx->stream);
        switch (mode) {
        case 0: ub_kernel<0><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 1: ub_kernel<1><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 2: ub_kernel<2><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 3: ub_kernel<3><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 4: ub_kernel<4><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 5: ub_kernel<5><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 6: ub_kernel<6><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 7: ub_kernel<7><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 8: ub_kernel<8><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 9: ub_kernel<9><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 10: ub_kernel<10><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 15: ub_kernel<15><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 16: ub_kernel<16><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        case 12: case 13: case 14: ub_fs_pattern_kernel<<<grid, nt, 0, ctx->stream>>>(iters / 16, d_seed, d_out); break;
        default: ub_kernel<11><<<grid, nt, 0, ctx->stream>>>(iters, d_seed, d_out); break;
        }
        ctx->launches++;
    }
    cudaEventRecord(e1, ctx->stream);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaGetLastError());
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_seed); cudaFree(d_out);
    const double per_step = mode >= 15 ? 32.0 : mode == 2 ? 12.0 : mode == 3 ? 16.0 : mode >= 12 ? 704.0 / 16.0 : 8.0;
    if (ms) *ms = t;
    if (px_ops) *px_ops = per_step * (double)iters * (double)grid * nt;
    return MEMCI_OK;
}
