// microbench.cu — B200 pipe-rate probes that ground the interact-kernel design.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu
// Run (GPU box): ./tools/microbench
//
// Questions answered:
//  1. FP32 FFMA peak (scalar)                  -> roofline denominator sanity (nominal 74.4 TFLOP/s @1965 MHz)
//  2. fma.rn.f32x2 (FFMA2) throughput          -> does the packed form halve issue slots at equal FLOP rate?
//  3. MUFU.RSQ throughput
//  4. prototype interaction inner loops (scalar vs packed, T targets/thread) -> cycles per interaction
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// ------------------------------------------------------------------ pure pipe probes
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, int iters, float a, float b) {
    float2 x[8];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 + independent ALU-pipe work (FSETP+FSEL), to see co-issue
__global__ void k_ffma2_alu(float* out, int iters, float a, float b) {
    float2 x[8];
    float y[8];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); y[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = __ffma2_rn(x[i], a2, b2);
            y[i] = (y[i] <= x[i].x) ? y[i] : b;   // FSETP + FSEL (data-dependent so it cannot be hoisted)
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mufu(float* out, int iters) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + 1.0f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------ interaction prototypes
constexpr int TS = 1024;  // sources per smem tile

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// scalar: one source at a time from AoS float4 smem, T targets per thread, inline mask d <= rs+rt
template <int T>
__global__ void __launch_bounds__(256) k_proto_scalar(const float4* __restrict__ src, const float4* __restrict__ tgt,
                                                      float2* __restrict__ out, int reps) {
    __shared__ float4 tile[TS];
    for (int i = threadIdx.x; i < TS; i += blockDim.x) tile[i] = src[i];
    __syncthreads();
    float tx[T], ty[T], tr[T], ax[T], ay[T];
#pragma unroll
    for (int k = 0; k < T; ++k) {
        float4 t = tgt[(blockIdx.x * blockDim.x + threadIdx.x) * T + k];
        tx[k] = t.x; ty[k] = t.y; tr[k] = t.w; ax[k] = 0; ay[k] = 0;
    }
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 4
        for (int j = 0; j < TS; ++j) {
            const float4 s = tile[j];
#pragma unroll
            for (int k = 0; k < T; ++k) {
                float dx = s.x - tx[k], dy = s.y - ty[k];
                float d2 = fmaf(dy, dy, dx * dx);
                float inv = rsqrt_approx(d2);
                float d = d2 * inv;
                float R = s.w + tr[k];
                float w = inv * inv * s.z;
                w = (d <= R) ? 0.f : w;
                ax[k] = fmaf(dx, w, ax[k]);
                ay[k] = fmaf(dy, w, ay[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < T; ++k) out[(blockIdx.x * blockDim.x + threadIdx.x) * T + k] = make_float2(ax[k], ay[k]);
}

// packed: two sources per f32x2 op, SoA smem {x[],y[],m[],r[]}, T targets per thread
template <int T, bool MASK>
__global__ void __launch_bounds__(256) k_proto_packed(const float4* __restrict__ src, const float4* __restrict__ tgt,
                                                      float2* __restrict__ out, int reps) {
    __shared__ __align__(16) float sx[TS], sy[TS], sm[TS], sr[TS];
    for (int i = threadIdx.x; i < TS; i += blockDim.x) { float4 s = src[i]; sx[i] = s.x; sy[i] = s.y; sm[i] = s.z; sr[i] = s.w; }
    __syncthreads();
    float2 ntx[T], nty[T], tr[T], ax[T], ay[T];
#pragma unroll
    for (int k = 0; k < T; ++k) {
        float4 t = tgt[(blockIdx.x * blockDim.x + threadIdx.x) * T + k];
        ntx[k] = make_float2(-t.x, -t.x); nty[k] = make_float2(-t.y, -t.y); tr[k] = make_float2(t.w, t.w);
        ax[k] = make_float2(0, 0); ay[k] = make_float2(0, 0);
    }
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
        for (int j = 0; j < TS; j += 4) {
            const float4 X = *reinterpret_cast<const float4*>(&sx[j]);
            const float4 Y = *reinterpret_cast<const float4*>(&sy[j]);
            const float4 M = *reinterpret_cast<const float4*>(&sm[j]);
            const float4 R = *reinterpret_cast<const float4*>(&sr[j]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 x2 = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                const float2 y2 = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                const float2 m2 = h ? make_float2(M.z, M.w) : make_float2(M.x, M.y);
                const float2 r2 = h ? make_float2(R.z, R.w) : make_float2(R.x, R.y);
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    float2 dx = __fadd2_rn(x2, ntx[k]);
                    float2 dy = __fadd2_rn(y2, nty[k]);
                    float2 d2 = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
                    float2 inv = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
                    float2 w = __fmul2_rn(__fmul2_rn(inv, inv), m2);
                    if (MASK) {
                        float2 d = __fmul2_rn(d2, inv);
                        float2 RR = __fadd2_rn(r2, tr[k]);
                        w.x = (d.x <= RR.x) ? 0.f : w.x;
                        w.y = (d.y <= RR.y) ? 0.f : w.y;
                    }
                    ax[k] = __ffma2_rn(dx, w, ax[k]);
                    ay[k] = __ffma2_rn(dy, w, ay[k]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < T; ++k)
        out[(blockIdx.x * blockDim.x + threadIdx.x) * T + k] = make_float2(ax[k].x + ax[k].y, ay[k].x + ay[k].y);
}

// ------------------------------------------------------------------ host
template <class F>
static float time_ms(F launch, int warm = 2, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < warm; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d MHz\n", prop.name, sms, prop.clockRate / 1000);
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256 * 8));
    const int iters = 1 << 14;

    for (int bps : {4, 8}) {
        float ms = time_ms([&] { k_ffma<<<sms * bps, 256>>>(out, iters, 1.0001f, 0.5f); });
        double fl = 2.0 * 16 * iters * 256.0 * sms * bps;
        printf("FFMA   scalar  %d blk/SM: %8.3f ms  %7.2f TFLOP/s  (%.3f instr/clk/SMSP @1965)\n", bps, ms, fl / ms * 1e-9,
               fl / 2 / 32 / (ms * 1e-3) / (sms * 4) / 1.965e9);
        ms = time_ms([&] { k_ffma2<<<sms * bps, 256>>>(out, iters, 1.0001f, 0.5f); });
        printf("FFMA2  packed  %d blk/SM: %8.3f ms  %7.2f TFLOP/s  (%.3f instr/clk/SMSP @1965)\n", bps, ms, fl / ms * 1e-9,
               fl / 4 / 32 / (ms * 1e-3) / (sms * 4) / 1.965e9);
        ms = time_ms([&] { k_ffma2_alu<<<sms * bps, 256>>>(out, iters, 1.0001f, 0.5f); });
        printf("FFMA2+FSETP+FSEL %d blk/SM: %8.3f ms  %7.2f TFLOP/s (fma part)\n", bps, ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_mufu<<<sms * bps, 256>>>(out, iters / 4); });
        double mu = 8.0 * (iters / 4) * 256.0 * sms * bps;
        printf("MUFU.RSQ       %d blk/SM: %8.3f ms  %7.2f Gop/s  (%.2f /clk/SM @1965)\n", bps, ms, mu / ms * 1e-6,
               mu / (ms * 1e-3) / sms / 1.965e9);
    }

    // interaction prototypes
    float4 *src, *tgt; float2* o2;
    const int maxT = 4, nblk = sms * 4;
    CK(cudaMalloc(&src, sizeof(float4) * TS));
    CK(cudaMalloc(&tgt, sizeof(float4) * nblk * 256 * maxT));
    CK(cudaMalloc(&o2, sizeof(float2) * nblk * 256 * maxT));
    {
        float4* h = (float4*)malloc(sizeof(float4) * nblk * 256 * maxT);
        srand(1);
        auto rnd = [] { return float(rand()) / RAND_MAX; };
        for (int i = 0; i < nblk * 256 * maxT; ++i) h[i] = make_float4((rnd() - .5f) * 1.5e10f, (rnd() - .5f) * 1.5e10f, 1e26f + rnd() * 3e27f, 3e7f + rnd() * 6e7f);
        CK(cudaMemcpy(tgt, h, sizeof(float4) * nblk * 256 * maxT, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(src, h + 12345, sizeof(float4) * TS, cudaMemcpyHostToDevice));
        free(h);
    }
    const int reps = 64;
    auto report = [&](const char* name, int T, int blocks, float ms) {
        double inter = double(blocks) * 256 * T * TS * reps;
        double rate = inter / (ms * 1e-3);
        printf("%-28s T=%d blocks=%4d: %8.3f ms  %.3e inter/s  %5.2f clk/inter/SMSP-warp @1965  (%.1f%% of 74.4T @18flop)\n", name, T, blocks, ms, rate,
               (sms * 4 * 32 * 1.965e9) / rate, rate * 18 / 74.4e12 * 100);
    };
    for (int bps : {2, 3, 4}) {
        int blocks = sms * bps;
        report("scalar mask", 1, blocks, time_ms([&] { k_proto_scalar<1><<<blocks, 256>>>(src, tgt, o2, reps); }));
        report("scalar mask", 2, blocks, time_ms([&] { k_proto_scalar<2><<<blocks, 256>>>(src, tgt, o2, reps); }));
        report("scalar mask", 4, blocks, time_ms([&] { k_proto_scalar<4><<<blocks, 256>>>(src, tgt, o2, reps); }));
        report("packed mask", 1, blocks, time_ms([&] { k_proto_packed<1, true><<<blocks, 256>>>(src, tgt, o2, reps); }));
        report("packed mask", 2, blocks, time_ms([&] { k_proto_packed<2, true><<<blocks, 256>>>(src, tgt, o2, reps); }));
        report("packed mask", 4, blocks, time_ms([&] { k_proto_packed<4, true><<<blocks, 256>>>(src, tgt, o2, reps); }));
        report("packed nomask", 2, blocks, time_ms([&] { k_proto_packed<2, false><<<blocks, 256>>>(src, tgt, o2, reps); }));
        report("packed nomask", 4, blocks, time_ms([&] { k_proto_packed<4, false><<<blocks, 256>>>(src, tgt, o2, reps); }));
    }
    printf("done\n");
    return 0;
}
