// Autoregressive-decode kernels (one new token per sequence per step).  At batch <= 16 every linear layer
// is a weight-streaming problem: the 7.86 GB of bf16 weights are read once per step and HBM bandwidth is
// the bound, so these are coalesced CUDA-core kernels, not tensor-core tiles.
//
//   cv_linear_small_m : y[M,N] = x[M,K] W[N,K]^T + b (+GELU) (+abs-max), M <= 16
//                       (F.linear of mpu/layers.py:243,319 as reached from the sampling loop,
//                        generation/sampling.py:147-151, and the last-token logits GEMM, model/gpt2_modeling.py:117)
//   cv_attn_decode    : one query per sequence against the K|V cache, with the new token's K/V appended
//                       in the same kernel (standard_attention, mpu/sparse_transformer.py:652-673, for sq = 1,
//                       where every cached key is visible)
#include "common.cuh"
#include "host.h"
#include "../../include/cogview_b200.h"

namespace {
using namespace cv;

// ------------------------------------------------------------------------------------------------
// skinny linear
// ------------------------------------------------------------------------------------------------
constexpr int SK_WARPS = 4;
constexpr int SK_COLS = 4;        // output columns per warp
constexpr int SK_KC = 2048;       // K chunk staged in shared memory

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        f[2 * t] = __low2float(p[t]);
        f[2 * t + 1] = __high2float(p[t]);
    }
}

template <int MT>
__global__ void __launch_bounds__(SK_WARPS * 32)
linear_small_m_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ W,
                      int64_t ldw, const __nv_bfloat16* __restrict__ bias, void* __restrict__ out, int64_t ldo,
                      int out_f32, int act, float* __restrict__ absmax, int M, int N, int K) {
    __shared__ __align__(16) __nv_bfloat16 xs[MT][SK_KC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = (blockIdx.x * SK_WARPS + warp) * SK_COLS;
    float acc[MT][SK_COLS];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int c = 0; c < SK_COLS; ++c) acc[m][c] = 0.f;

    for (int kc = 0; kc < K; kc += SK_KC) {
        const int klen = min(SK_KC, K - kc);
        __syncthreads();
        for (int i = threadIdx.x * 8; i < MT * SK_KC; i += SK_WARPS * 32 * 8) {
            const int m = i / SK_KC, k = i % SK_KC;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (m < M && k < klen) v = *reinterpret_cast<const uint4*>(x + (size_t)m * ldx + kc + k);
            *reinterpret_cast<uint4*>(&xs[m][k]) = v;
        }
        __syncthreads();
        if (n0 < N) {
            for (int k = lane * 8; k < klen; k += 256) {
                float w[SK_COLS][8];
#pragma unroll
                for (int c = 0; c < SK_COLS; ++c) {
                    uint4 u = make_uint4(0, 0, 0, 0);
                    if (n0 + c < N) u = __ldg(reinterpret_cast<const uint4*>(W + (size_t)(n0 + c) * ldw + kc + k));
                    bf16x8_to_float(u, w[c]);
                }
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    float xv[8];
                    bf16x8_to_float(*reinterpret_cast<const uint4*>(&xs[m][k]), xv);
#pragma unroll
                    for (int c = 0; c < SK_COLS; ++c) {
#pragma unroll
                        for (int t = 0; t < 8; ++t) acc[m][c] = fmaf(w[c][t], xv[t], acc[m][c]);
                    }
                }
            }
        }
    }
    if (n0 >= N) return;
    float tmax = 0.f;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int c = 0; c < SK_COLS; ++c) {
            float v = warp_sum(acc[m][c]);
            if (lane == m * SK_COLS + c && m < M && n0 + c < N) {
                if (bias != nullptr) v += __bfloat162float(bias[n0 + c]);
                if (act == 1) v = gelu_tanh(v);
                if (out_f32) {
                    static_cast<float*>(out)[(size_t)m * ldo + n0 + c] = v;
                    tmax = fabsf(v);
                } else {
                    __nv_bfloat16 o = __float2bfloat16_rn(v);
                    static_cast<__nv_bfloat16*>(out)[(size_t)m * ldo + n0 + c] = o;
                    tmax = fabsf(__bfloat162float(o));
                }
            }
        }
    }
    if (absmax != nullptr) {
        tmax = warp_max(tmax);
        if (lane == 0 && tmax > 0.f) atomic_max_nonneg(absmax, tmax);
    }
}

// ------------------------------------------------------------------------------------------------
// decode attention
// ------------------------------------------------------------------------------------------------
constexpr int DA_WARPS = 4;
constexpr int HD = 64;

struct DecodeParams {
    const __nv_bfloat16* qkv;   // [b, 3h]: q | k_new | v_new of the token at position cur_len
    __nv_bfloat16* cache;       // [b, max_len, 2h]: K | V
    int64_t cache_bs;           // batch stride (elements)
    const int* cur_len_dev;     // device int32: number of cached tokens BEFORE this step (or null -> cur_len)
    int cur_len;
    __nv_bfloat16* out;         // [b, h]
    float* partial;             // [b, heads, nsplit, HD + 2] when nsplit > 1
    int heads, nsplit, max_len;
    float scale_log2;
};

// 8 lanes cooperate on one key (8 dims each); a warp covers 4 keys per iteration.
__global__ void __launch_bounds__(DA_WARPS * 32)
attn_decode_kernel(const DecodeParams p) {
    __shared__ float s_m[DA_WARPS * 4], s_l[DA_WARPS * 4];
    __shared__ float s_acc[DA_WARPS * 4][HD];
    const int head = blockIdx.x, batch = blockIdx.y, split = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, sub = lane & 7;
    const int h = p.heads * HD;
    const int t = p.cur_len_dev ? *p.cur_len_dev : p.cur_len;   // cached tokens; the new token sits at index t
    const __nv_bfloat16* qrow = p.qkv + (size_t)batch * 3 * h + head * HD + sub * 8;
    float q[8], kn[8], vn[8];
    bf16x8_to_float(*reinterpret_cast<const uint4*>(qrow), q);
    const uint4 knew = *reinterpret_cast<const uint4*>(qrow + h);
    const uint4 vnew = *reinterpret_cast<const uint4*>(qrow + 2 * h);
    bf16x8_to_float(knew, kn);
    bf16x8_to_float(vnew, vn);
    __nv_bfloat16* kbase = p.cache + (size_t)batch * p.cache_bs + head * HD + sub * 8;
    // append (one split does it; the values are also used straight from registers below)
    if (split == p.nsplit - 1 && warp == 0 && grp == 0 && t < p.max_len) {
        *reinterpret_cast<uint4*>(kbase + (size_t)t * 2 * h) = knew;
        *reinterpret_cast<uint4*>(kbase + (size_t)t * 2 * h + h) = vnew;
    }
    const int total = t + 1;                                    // keys 0..t
    const int per = (total + p.nsplit - 1) / p.nsplit;
    const int j0 = split * per, j1 = min(total, j0 + per);
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // the trip count is warp-uniform (the shuffles below need all 32 lanes); out-of-range keys are skipped
    for (int jb = j0 + warp * 4; jb < j1; jb += DA_WARPS * 4) {
        const int j = jb + grp;
        const bool valid = j < j1;
        float kf[8], vf[8];
        if (!valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { kf[i] = 0.f; vf[i] = 0.f; }
        } else if (j == t) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { kf[i] = kn[i]; vf[i] = vn[i]; }
        } else {
            const __nv_bfloat16* kp = kbase + (size_t)j * 2 * h;
            bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(kp)), kf);
            bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(kp + h)), vf);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(q[i], kf[i], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (valid) {
            s *= p.scale_log2;
            const float mn = fmaxf(m, s);
            const float alpha = exp2f(m - mn), pr = exp2f(s - mn);
            m = mn;
            l = l * alpha + pr;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = acc[i] * alpha + pr * vf[i];
        }
    }
    // combine the 16 (warp, group) partial states through shared memory
    const int slot = warp * 4 + grp;
    if (sub == 0) { s_m[slot] = m; s_l[slot] = l; }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[slot][sub * 8 + i] = acc[i];
    __syncthreads();
    if (threadIdx.x < HD) {
        const int d = threadIdx.x;
        float M = -INFINITY;
        for (int sidx = 0; sidx < DA_WARPS * 4; ++sidx) M = fmaxf(M, s_m[sidx]);
        float L = 0.f, A = 0.f;
        for (int sidx = 0; sidx < DA_WARPS * 4; ++sidx) {
            const float w = (s_m[sidx] == -INFINITY) ? 0.f : exp2f(s_m[sidx] - M);
            L += s_l[sidx] * w;
            A += s_acc[sidx][d] * w;
        }
        if (p.nsplit == 1) {
            p.out[(size_t)batch * h + head * HD + d] = __float2bfloat16_rn(A / L);
        } else {
            float* dst = p.partial + (((size_t)batch * p.heads + head) * p.nsplit + split) * (HD + 2);
            dst[d] = A;
            if (d == 0) { dst[HD] = M; dst[HD + 1] = L; }
        }
    }
}

__global__ void attn_decode_combine_kernel(const float* __restrict__ partial, __nv_bfloat16* __restrict__ out,
                                           int heads, int nsplit) {
    const int head = blockIdx.x, batch = blockIdx.y, d = threadIdx.x;
    const float* src = partial + ((size_t)batch * heads + head) * nsplit * (HD + 2);
    float M = -INFINITY;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, src[s * (HD + 2) + HD]);
    float L = 0.f, A = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        const float ms = src[s * (HD + 2) + HD];
        const float w = (ms == -INFINITY) ? 0.f : exp2f(ms - M);
        L += src[s * (HD + 2) + HD + 1] * w;
        A += src[s * (HD + 2) + d] * w;
    }
    out[(size_t)batch * heads * HD + head * HD + d] = __float2bfloat16_rn(A / L);
}

}  // namespace

extern "C" int cv_linear_small_m(const void* x, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* out,
                                 int64_t ldo, int out_is_f32, int act, float* absmax, int M, int N, int K,
                                 void* stream) {
    CV_REQUIRE(x && W && out, "null pointer");
    CV_REQUIRE(M >= 1 && M <= 16, "cv_linear_small_m handles 1 <= M <= 16 rows (use cv_gemm_bf16 above that)");
    CV_REQUIRE(N > 0 && K > 0 && K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "K, ldx, ldw must be multiples of 8");
    CV_REQUIRE(act == 0 || act == 1, "act must be 0 or 1");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int cols_per_block = SK_WARPS * SK_COLS;
    const int grid = (N + cols_per_block - 1) / cols_per_block;
    const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
    const __nv_bfloat16* wb = static_cast<const __nv_bfloat16*>(W);
    const __nv_bfloat16* bb = static_cast<const __nv_bfloat16*>(bias);
#define LAUNCH(MT) linear_small_m_kernel<MT><<<grid, SK_WARPS * 32, 0, s>>>(xb, ldx, wb, ldw, bb, out, ldo, out_is_f32, act, absmax, M, N, K)
    if (M == 1) LAUNCH(1);
    else if (M == 2) LAUNCH(2);
    else if (M <= 4) LAUNCH(4);
    else if (M <= 8) LAUNCH(8);
    else {
        // 16 rows x 2048 x 2 B = 64 KB of static smem would exceed the 48 KB static limit: two passes of 8 rows
        LAUNCH(8);
        CV_LAUNCH_CHECK();
        const int M2 = M - 8;
        const size_t osz = out_is_f32 ? 4 : 2;
        linear_small_m_kernel<8><<<grid, SK_WARPS * 32, 0, s>>>(xb + 8 * ldx, ldx, wb, ldw, bb,
                                                              static_cast<char*>(out) + 8 * ldo * osz, ldo,
                                                              out_is_f32, act, absmax, M2, N, K);
    }
#undef LAUNCH
    CV_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t cv_attn_decode_workspace_bytes(int b, int heads, int nsplit) {
    return (int64_t)b * heads * nsplit * (HD + 2) * sizeof(float);
}

extern "C" int cv_attn_decode(const void* qkv, void* cache, int64_t cache_batch_stride, const int* cur_len_dev,
                              int cur_len, void* out, float* workspace, int b, int heads, int head_dim, int max_len,
                              int nsplit, void* stream) {
    CV_REQUIRE(qkv && cache && out, "null pointer");
    CV_REQUIRE(head_dim == HD, "head_dim must be 64");
    CV_REQUIRE(b > 0 && heads > 0 && max_len > 0 && nsplit >= 1 && nsplit <= 64, "bad sizes");
    CV_REQUIRE(nsplit == 1 || workspace != nullptr, "workspace required when nsplit > 1");
    CV_REQUIRE(cur_len_dev != nullptr || (cur_len >= 0 && cur_len < max_len), "cur_len out of range");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DecodeParams p;
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.cache = static_cast<__nv_bfloat16*>(cache);
    p.cache_bs = cache_batch_stride;
    p.cur_len_dev = cur_len_dev;
    p.cur_len = cur_len;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.partial = workspace;
    p.heads = heads; p.nsplit = nsplit; p.max_len = max_len;
    p.scale_log2 = (1.0f / sqrtf((float)head_dim)) * 1.4426950408889634f;
    dim3 grid(heads, b, nsplit);
    attn_decode_kernel<<<grid, DA_WARPS * 32, 0, s>>>(p);
    CV_LAUNCH_CHECK();
    if (nsplit > 1) {
        attn_decode_combine_kernel<<<dim3(heads, b), HD, 0, s>>>(workspace, static_cast<__nv_bfloat16*>(out), heads,
                                                                nsplit);
        CV_LAUNCH_CHECK();
    }
    return 0;
}
