// microbench2.cu — second round of pipe probes: where do FSETP/FSEL/MUFU go, and which mask formulation is cheapest.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- pipe probes: N_FMA2 packed fmas + N_MUFU rsqrt + N_SETP (fsetp+fsel pairs) + N_MNMX fmnmx per iteration, all independent chains
template <int N_FMA2, int N_MUFU, int N_SEL, int N_MNMX>
__global__ void __launch_bounds__(256) k_mix(float* out, int iters, float a, float b) {
    float2 x[8]; float m[8]; float y[8]; float z[8];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); m[i] = threadIdx.x + 1.f + i; y[i] = i; z[i] = i * 3.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N_FMA2; ++i) x[i % 8] = __ffma2_rn(x[i % 8], a2, b2);
#pragma unroll
        for (int i = 0; i < N_MUFU; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i % 8]));
#pragma unroll
        for (int i = 0; i < N_SEL; ++i) y[i % 8] = (y[i % 8] <= z[i % 8]) ? z[(i + 1) % 8] : y[i % 8];
#pragma unroll
        for (int i = 0; i < N_MNMX; ++i) z[i % 8] = fminf(z[i % 8], y[(i + 3) % 8]) + 0.f * a;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + m[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>  // 0 FADD2, 1 FMUL2, 2 scalar FADD, 3 scalar FMUL
__global__ void __launch_bounds__(256) k_op(float* out, int iters, float a) {
    float2 x[8]; const float2 a2 = make_float2(a, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) x[i] = __fadd2_rn(x[i], a2);
            if (OP == 1) x[i] = __fmul2_rn(x[i], a2);
            if (OP == 2) { x[i].x = __fadd_rn(x[i].x, a); x[i].y = __fadd_rn(x[i].y, a); }
            if (OP == 3) { x[i].x = __fmul_rn(x[i].x, a); x[i].y = __fmul_rn(x[i].y, a); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- interaction prototypes
constexpr int TS = 1024;
// VAR: 1 = rsqrt + d<=R mask (FSEL)          [round-1 'packed mask']
//      2 = rcp   + d2<=R^2 mask (FSEL)
//      3 = rcp   + conservative per-target d2<=thr, OR-accumulated flag, no FSEL (flag -> branch per 4-source group)
//      4 = V2 with one rcp per two interactions
//      5 = V3 with one rcp per two interactions
//      6 = rcp   + FSEL mask on conservative per-target thr (no R add, no R^2)
template <int T, int VAR>
__global__ void __launch_bounds__(256) k_proto(const float4* __restrict__ src, const float4* __restrict__ tgt,
                                               float2* __restrict__ out, int* __restrict__ flags, int reps) {
    __shared__ __align__(16) float sx[TS], sy[TS], sm[TS], sr[TS];
    for (int i = threadIdx.x; i < TS; i += blockDim.x) { float4 s = src[i]; sx[i] = s.x; sy[i] = s.y; sm[i] = s.z; sr[i] = s.w; }
    __syncthreads();
    float tx[T], ty[T], tr[T], thr[T]; float2 ax[T], ay[T];
    int nflag = 0;
#pragma unroll
    for (int k = 0; k < T; ++k) {
        float4 t = tgt[(blockIdx.x * blockDim.x + threadIdx.x) * T + k];
        tx[k] = t.x; ty[k] = t.y; tr[k] = t.w; thr[k] = (t.w + 9e7f) * (t.w + 9e7f);
        ax[k] = make_float2(0, 0); ay[k] = make_float2(0, 0);
    }
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
        for (int j = 0; j < TS; j += 4) {
            const float4 X = *reinterpret_cast<const float4*>(&sx[j]);
            const float4 Y = *reinterpret_cast<const float4*>(&sy[j]);
            const float4 M = *reinterpret_cast<const float4*>(&sm[j]);
            float4 R;
            if (VAR == 1 || VAR == 2 || VAR == 4) R = *reinterpret_cast<const float4*>(&sr[j]);
            bool any = false;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 x2 = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                const float2 y2 = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                const float2 m2 = h ? make_float2(M.z, M.w) : make_float2(M.x, M.y);
                const float2 r2 = h ? make_float2(R.z, R.w) : make_float2(R.x, R.y);
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    float2 dx = __fadd2_rn(x2, make_float2(-tx[k], -tx[k]));
                    float2 dy = __fadd2_rn(y2, make_float2(-ty[k], -ty[k]));
                    float2 d2 = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
                    float2 w;
                    if (VAR == 1) {
                        float2 inv = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
                        w = __fmul2_rn(__fmul2_rn(inv, m2), inv);
                        float2 d = __fmul2_rn(d2, inv);
                        float2 RR = __fadd2_rn(r2, make_float2(tr[k], tr[k]));
                        w.x = (d.x <= RR.x) ? 0.f : w.x;
                        w.y = (d.y <= RR.y) ? 0.f : w.y;
                    } else {
                        float2 rc;
                        if (VAR == 4 || VAR == 5) {
                            float q = rcp_approx(d2.x * d2.y);
                            rc = make_float2(q * d2.y, q * d2.x);
                        } else {
                            rc = make_float2(rcp_approx(d2.x), rcp_approx(d2.y));
                        }
                        w = __fmul2_rn(rc, m2);
                        if (VAR == 2 || VAR == 4) {
                            float2 RR = __fadd2_rn(r2, make_float2(tr[k], tr[k]));
                            RR = __fmul2_rn(RR, RR);
                            w.x = (d2.x <= RR.x) ? 0.f : w.x;
                            w.y = (d2.y <= RR.y) ? 0.f : w.y;
                        } else if (VAR == 6) {
                            w.x = (d2.x <= thr[k]) ? 0.f : w.x;
                            w.y = (d2.y <= thr[k]) ? 0.f : w.y;
                        } else {
                            any = any || (d2.x <= thr[k]) || (d2.y <= thr[k]);
                        }
                    }
                    ax[k] = __ffma2_rn(dx, w, ax[k]);
                    ay[k] = __ffma2_rn(dy, w, ay[k]);
                }
            }
            if (VAR == 3 || VAR == 5) { if (any) { ++nflag; } }
        }
    }
#pragma unroll
    for (int k = 0; k < T; ++k)
        out[(blockIdx.x * blockDim.x + threadIdx.x) * T + k] = make_float2(ax[k].x + ax[k].y, ay[k].x + ay[k].y);
    if (nflag) atomicAdd(flags, nflag);
}

template <class F>
static float time_ms(F launch, int warm = 2, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < warm; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256 * 8));
    const int iters = 1 << 13, bps = 4;
    const double wi = double(iters) * 8 * sms * bps;  // warp-iterations
    auto cyc = [&](float ms) { return ms * 1e-3 * 1.965e9 * sms * 4 / wi; };  // SMSP cycles per warp-iteration
#define MIX(a, b, c, d) printf("mix FFMA2=%2d MUFU=%d SETP+SEL=%2d FMNMX+FADD=%2d : %6.2f cyc/iter\n", a, b, c, d, cyc(time_ms([&] { k_mix<a, b, c, d><<<sms * bps, 256>>>(out, iters, 1.0001f, 0.5f); })))
    MIX(8, 0, 0, 0); MIX(0, 2, 0, 0); MIX(0, 0, 8, 0); MIX(0, 0, 0, 8);
    MIX(8, 2, 0, 0); MIX(8, 0, 4, 0); MIX(8, 0, 8, 0); MIX(8, 2, 4, 0); MIX(10, 2, 4, 0); MIX(8, 2, 2, 0); MIX(7, 2, 2, 0); MIX(8, 1, 4, 0); MIX(9, 1, 4, 0); MIX(4, 2, 0, 0); MIX(0, 2, 8, 0); MIX(0, 2, 4, 0);
#define OPB(n, name) printf("%s: %6.2f cyc / 8 ops\n", name, cyc(time_ms([&] { k_op<n><<<sms * bps, 256>>>(out, iters, 1.0001f); })))
    OPB(0, "FADD2"); OPB(1, "FMUL2"); OPB(2, "FADD x2 scalar"); OPB(3, "FMUL x2 scalar");

    float4 *src, *tgt; float2* o2; int* flags;
    const int maxT = 4, nblk = sms * 4;
    CK(cudaMalloc(&src, sizeof(float4) * TS)); CK(cudaMalloc(&flags, 4)); CK(cudaMemset(flags, 0, 4));
    CK(cudaMalloc(&tgt, sizeof(float4) * nblk * 256 * maxT)); CK(cudaMalloc(&o2, sizeof(float2) * nblk * 256 * maxT));
    {
        float4* h = (float4*)malloc(sizeof(float4) * nblk * 256 * maxT);
        srand(1); auto rnd = [] { return float(rand()) / RAND_MAX; };
        for (int i = 0; i < nblk * 256 * maxT; ++i) h[i] = make_float4((rnd() - .5f) * 15.f, (rnd() - .5f) * 15.f, 1e26f + rnd() * 3e27f, 0.03f + rnd() * 0.06f);
        CK(cudaMemcpy(tgt, h, sizeof(float4) * nblk * 256 * maxT, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(src, h + 12345, sizeof(float4) * TS, cudaMemcpyHostToDevice)); free(h);
    }
    const int reps = 64;
    auto report = [&](const char* name, int T, int blocks, float ms) {
        double inter = double(blocks) * 256 * T * TS * reps; double rate = inter / (ms * 1e-3);
        printf("%-34s T=%d blocks=%4d: %8.3f ms  %.3e inter/s  %5.2f clk/inter  (%.1f%% of 74.4T @18flop)\n", name, T, blocks, ms, rate,
               (sms * 4 * 32 * 1.965e9) / rate, rate * 18 / 74.4e12 * 100);
    };
    int blocks = sms * 4;
#define PR(T, V, name) report(name, T, blocks, time_ms([&] { k_proto<T, V><<<blocks, 256>>>(src, tgt, o2, flags, reps); }))
    PR(2, 1, "V1 rsqrt d<=R FSEL"); PR(4, 1, "V1 rsqrt d<=R FSEL");
    PR(2, 2, "V2 rcp d2<=R2 FSEL"); PR(4, 2, "V2 rcp d2<=R2 FSEL");
    PR(2, 3, "V3 rcp cons-thr OR-flag"); PR(4, 3, "V3 rcp cons-thr OR-flag");
    PR(2, 4, "V4 V2 + shared rcp"); PR(4, 4, "V4 V2 + shared rcp");
    PR(2, 5, "V5 V3 + shared rcp"); PR(4, 5, "V5 V3 + shared rcp");
    PR(2, 6, "V6 rcp cons-thr FSEL"); PR(4, 6, "V6 rcp cons-thr FSEL");
    int hf; CK(cudaMemcpy(&hf, flags, 4, cudaMemcpyDeviceToHost)); printf("flags %d\ndone\n", hf);
    return 0;
}
