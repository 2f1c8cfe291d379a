// One autoregressive decode step (one new token per sequence, all layers + logits) as ONE persistent kernel.
//
// Reference path: generation/sampling.py:147-151 calling GPT2Model.forward (model/gpt2_modeling.py:106-123) with one
// token per sequence: per layer  LN1 -> QKV linear (mpu/layers.py:243) -> attention over the memory
// (mpu/sparse_transformer.py:652-673) -> dense (mpu/layers.py:319) -> x + LN3 -> LN2 -> h->4h + GELU -> 4h->h ->
// y + LN4 (mpu/sparse_transformer.py:314-342), then the final LayerNorm and the tied-embedding logits.
//
// The step is HBM bound: 7.86 GB of bf16 weights are read once per step.  Launching one small kernel
// per linear leaves every launch in its ramp-up / drain (14 us for a 6 us-ideal launch, profiles/r01_*linear*), so:
//
//   * one CTA per SM, resident for the whole step (cooperative launch): 8 consumer warps + 1 producer warp;
//   * every weight matrix is split into contiguous, byte-balanced row ranges, one per CTA (the 4h->h matrix as
//     37 row ranges x 4 K quarters so that the activation operand of every matrix is [M, h]);
//   * the producer warp walks the CTA's static schedule (all matrices of all layers, then the vocabulary matrix)
//     and copies 16-row x KS-column slabs into a shared-memory ring with cp.async.bulk (one bulk copy per weight
//     row segment, >= 512 B contiguous, L2 evict-first), completion on an mbarrier per stage.  Weights do not
//     depend on activations, so the producer never waits for a grid barrier: while the consumers sit in one of the
//     5 grid barriers of a layer (or in the attention phase) the ring fills with the next matrix and the HBM
//     stream never stops;
//   * consumer warps split K inside a stage, feed mma.sync.m16n8k16 (weights = A, 16 output columns as rows;
//     activations = B, up to 8 sequences as columns) from shared memory with conflict-free 16-byte loads (row pitch
//     = 64 mod 128 bytes; the k-permutation of the fragments is the same for A and B, so a dot product is unchanged),
//     reduce the 8 partial tiles through shared memory and apply bias / GELU;
//   * the Sandwich-LN glue (two abs-max LayerNorms + residual, mpu/sparse_transformer.py:40-44) is computed
//     redundantly by every CTA straight into its shared-memory activation operand — no single-CTA kernels between
//     the linears; the fp32 residual stream lives in two L2-resident buffers (owner CTAs write their slice);
//   * attention over the K|V cache runs as (sequence, head, key-range) units, one per consumer warp; the last unit
//     of a (sequence, head) to finish merges the partial softmax states (arrival counter), which avoids a sixth grid
//     barrier; the new token's K/V are appended in place;
//   * the K quarters of the 4h->h product are merged the same way (fixed summation order: deterministic).
//
// Everything exchanged between CTAs goes through L2 (ld.global.cg / st + fence + grid barrier): L1 is not coherent
// and there is no kernel boundary to invalidate it.
#include <type_traits>

#include "common.cuh"
#include "host.h"
#include "../../include/cogview_b200.h"

namespace {
using namespace cv;
typedef __nv_bfloat16 bf16;

constexpr int CW = 16;              // consumer warps: two groups of GW that take alternate ring stages
constexpr int GW = 8;               // warps per group = K slices of a stage
constexpr int CT = CW * 32;         // consumer threads
constexpr int NT = CT + 32;         // + one producer warp
constexpr int TILE = 16;            // weight rows (output columns) per MMA tile
constexpr int KG = 4;               // K groups of the 4h->h matrix
constexpr int HD = 64;              // head dim
constexpr int MAXST = 8;            // ring stages (upper bound)
constexpr int MAXM = 8;
constexpr int KVB = 64;             // keys per K|V ring stage (16 warps x 4 keys)
constexpr int KVB_MAX = 64;         // key blocks per (sequence, head): max_len <= KVB * KVB_MAX = 4096
constexpr int PART_STRIDE = HD + 2; // attention partial: acc[64], m, l

struct Params {
    const cv_decode_layer* layers;
    int L, h, heads, V, M, max_len;
    float eps, eps_final;
    const bf16 *wte, *wpe, *lnf_g, *lnf_b;
    const int64_t *ids, *pos;
    const int* cur_len;
    bf16* cache;
    int64_t cache_ls, cache_bs;
    float* logits;
    int64_t ldl;
    // workspace
    bf16 *qkv, *ctx, *attn_out, *h4, *mlp_out;
    float *resid_a, *resid_b, *attn_part;
    long long* fc2_acc;            // [2][MAXM][h] fixed-point (2^-40) sums of the 4h->h K quarters, see EPI_FIX64
    unsigned int* attn_cnt;
    unsigned long long *bar_ctr, *bar_base;
    int* err;
    unsigned long long* prof;      // optional [grid][L][16] globaltimer stamps of the phase boundaries
    // derived on the host
    int kstage, nst, stage_bytes, pitch, xpitch, S, quiet, pf_stages;
    float scale_log2;
};

// shared-memory carve-up (bytes from the 1024-aligned base)
constexpr int SM_BAR = 0;                                  // full[MAXST], empty[MAXST]
constexpr int SM_FLAG = 2 * MAXST * 8;                     // int flags
constexpr int SM_RED = 256;                                // float2 red[4][CW] (fits the [4][CW][MAXM + 1] slot)
constexpr int SM_PART = SM_RED + 4 * CW * (MAXM + 1) * 4;  // float part[2][CW][TILE][8]
constexpr int SM_XOP = ((SM_PART + 2 * CW * TILE * 8 * 4) + 127) / 128 * 128;

__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_u32(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                                int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bflo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bfhi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]) {
    f[0] = bflo(u.x); f[1] = bfhi(u.x); f[2] = bflo(u.y); f[3] = bfhi(u.y);
    f[4] = bflo(u.z); f[5] = bfhi(u.z); f[6] = bflo(u.w); f[7] = bfhi(u.w);
}
__device__ __forceinline__ uint32_t ldcg_u32(const void* p) { return __ldcg(reinterpret_cast<const unsigned int*>(p)); }
__device__ __forceinline__ uint4 ldcg_u128(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

__device__ __noinline__ void step_fail(const Params& p, int code) {
    if (p.err != nullptr) atomicExch(p.err, code);
    printf("cogview_b200: decode_step_kernel wait timed out (code %d, block %d, thread %d)\n", code, blockIdx.x,
           threadIdx.x);
    __trap();
}

// a CTA's share of one weight matrix: rows [r0, r1) x columns [k0, k0 + h)
struct Mat {
    const bf16* W;
    int64_t ldw;
    int r0, r1, k0;
};
__device__ __forceinline__ Mat make_mat(const void* W, int64_t ldw, int N, int part, int nparts, int k0) {
    Mat m;
    m.W = static_cast<const bf16*>(W);
    m.ldw = ldw;
    m.r0 = (int)(((int64_t)N * part) / nparts);
    m.r1 = (int)(((int64_t)N * (part + 1)) / nparts);
    m.k0 = k0;
    return m;
}

// EPI_FIX64: the four K quarters of the 4h->h product are summed by red.global.add.u64 on 2^-40 fixed-point images of
// the fp32 partial sums: integer addition is associative, so the result does not depend on the arrival order (a float
// atomic would make the step non-reproducible) and no merge pass / arrival counter is needed.
enum { EPI_BF16 = 0, EPI_BF16_GELU = 1, EPI_F32 = 2, EPI_FIX64 = 3 };
constexpr float FIX_SCALE = 1099511627776.0f;          // 2^40
constexpr float FIX_INV = 1.0f / 1099511627776.0f;

template <int MR, int CPW>
__global__ void __launch_bounds__(NT, 1) decode_step_kernel(const __grid_constant__ Params p,
                                                            const __grid_constant__ CUtensorMap tmKV) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, G = gridDim.x;
    const int h = p.h, M = p.M;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t* empty = full + MAXST;
    int* flags = reinterpret_cast<int*>(smem + SM_FLAG);
    float* red = reinterpret_cast<float*>(smem + SM_RED);
    float* part = reinterpret_cast<float*>(smem + SM_PART);
    const uint32_t xop = smem_u32(smem + SM_XOP);
    const uint32_t ring = xop + ((MAXM * p.xpitch + 127) / 128) * 128;
    const int nks = h / p.kstage;

    if (tid == 0) {
        flags[1] = p.quiet;
        for (int i = 0; i < p.nst; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], CW);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // ============================================================================================
    // producer warp: the weight stream
    // ============================================================================================
    if (warp == CW) {
        const uint64_t pol = evict_first_policy();
        const int t_p = __ldg(p.cur_len);
        const int nblk_p = (t_p + KVB - 1) / KVB;
        const long long NBp = (long long)p.M * p.heads * nblk_p;
        const int kf0 = (int)((NBp * cta) / G), kf1 = (int)((NBp * (cta + 1)) / G);
        // The CTA's static schedule as a cursor over ring stages: per layer the QKV matrix, the CTA's K|V key blocks,
        // dense, h->4h, 4h->h; then the vocabulary matrix.  Two cursors walk it: `ld` feeds the ring, `pf` runs
        // pf_stages further ahead and only pulls the bytes into L2 (cp.async.bulk.prefetch.L2) — when the consumers
        // come out of a grid barrier / glue / attention gap longer than the ring covers, they catch up from L2 (2-3x
        // the HBM rate) while the HBM stream itself never stopped.
        struct Cur {
            int it, r, ks, f;
            bool kv, done;
            Mat m;
        };
        auto set_mat = [&](Cur& c) {
            const int l = c.it >> 2, which = c.it & 3;
            if (c.it >= 4 * p.L) {
                c.m = make_mat(p.wte, h, p.V, cta, G, 0);
            } else {
                const cv_decode_layer& Lw = p.layers[l];
                if (which == 0) c.m = make_mat(Lw.w_qkv, h, 3 * h, cta, G, 0);
                else if (which == 1) c.m = make_mat(Lw.w_dense, h, h, cta, G, 0);
                else if (which == 2) c.m = make_mat(Lw.w_fc1, h, 4 * h, cta, G, 0);
                else c.m = make_mat(Lw.w_fc2, 4 * (int64_t)h, h, cta / KG, G / KG, (cta % KG) * h);
            }
            c.r = c.m.r0;
            c.ks = 0;
        };
        auto normalize = [&](Cur& c) {                     // point at an existing stage, or done
            while (!c.done) {
                if (c.kv) {
                    if (c.f < kf1) return;
                    c.kv = false;
                    ++c.it;
                    set_mat(c);
                } else if (c.r < c.m.r1) {
                    return;
                } else if (c.it < 4 * p.L && (c.it & 3) == 0) {
                    c.kv = true;                           // K|V blocks of this layer follow its QKV weights
                    c.f = kf0;
                } else if (c.it >= 4 * p.L) {
                    c.done = true;
                } else {
                    ++c.it;
                    set_mat(c);
                }
            }
        };
        auto step = [&](Cur& c) {
            if (c.kv) {
                ++c.f;
            } else if (++c.ks == nks) {
                c.ks = 0;
                c.r += TILE;
            }
            normalize(c);
        };
        auto issue = [&](const Cur& c, bool prefetch, int slot) {
            if (c.kv) {
                const int bh = c.f / nblk_p, blk = c.f - bh * nblk_p;
                const int head = bh % p.heads, batch = bh / p.heads, l = c.it >> 2;
                if (lane == 0) {
                    if (prefetch) {
                        tma_prefetch_4d(&tmKV, head * HD, blk * KVB, batch, l);
                        tma_prefetch_4d(&tmKV, h + head * HD, blk * KVB, batch, l);
                    } else {
                        mbar_expect_tx(&full[slot], 2 * KVB * 128);
                        const uint32_t dst = ring + slot * p.stage_bytes;
                        tma_load_4d_u32(dst, &tmKV, smem_u32(&full[slot]), head * HD, blk * KVB, batch, l);
                        tma_load_4d_u32(dst + KVB * 128, &tmKV, smem_u32(&full[slot]), h + head * HD, blk * KVB, batch, l);
                    }
                }
                __syncwarp();
            } else {
                const int rows = min(TILE, c.m.r1 - c.r);
                const bf16* src = c.m.W + (size_t)(c.r + lane) * c.m.ldw + c.m.k0 + c.ks * p.kstage;
                if (prefetch) {
                    if (lane < rows) bulk_prefetch_l2(src, (uint32_t)(p.kstage * 2));
                } else {
                    if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)(rows * p.kstage * 2));
                    __syncwarp();
                    if (lane < rows)
                        bulk_g2s(ring + slot * p.stage_bytes + lane * p.pitch, src, (uint32_t)(p.kstage * 2),
                                 smem_u32(&full[slot]), pol);
                }
            }
        };
        Cur ld, pf;
        ld.it = 0; ld.f = 0; ld.kv = false; ld.done = false;
        set_mat(ld);
        normalize(ld);
        pf = ld;
        int ahead = 0;                                     // stages between pf and ld
        const int ahead_min = p.nst, ahead_max = p.nst + p.pf_stages;
        if (p.pf_stages > 0)
            for (; ahead < ahead_min && !pf.done; ++ahead) step(pf);    // the ring itself covers the first nst stages
        int st = 0;
        uint32_t ph = 0;
        while (!ld.done) {
            if (p.pf_stages > 0) {
                for (; ahead < ahead_max && !pf.done; ++ahead) {
                    issue(pf, true, 0);
                    step(pf);
                }
            }
            mbar_wait(&empty[st], ph ^ 1);
            if (p.quiet) {      // experiment: no weight traffic while the consumers are in a latency phase
                while (*reinterpret_cast<volatile int*>(&flags[1]) != 0) __nanosleep(64);
            }
            issue(ld, false, st);
            step(ld);
            --ahead;
            if (++st == p.nst) { st = 0; ph ^= 1; }
        }
        return;
    }

    // ============================================================================================
    // consumer warps
    //
    // Every heavy routine below has exactly ONE call site, inside a phase loop: inlined copies of straight-line
    // code made the first version of this kernel 240 KB of SASS and instruction-cache bound (ncu: 64 % I-cache
    // hit rate, 17 % of the stall samples "no instruction"); one layer's working set now stays cache resident.
    // ============================================================================================
    const int g = lane >> 2, q = lane & 3;
    const int kpw = p.kstage / GW;                 // k elements per warp and stage ( = 32 * CPW )
    const int grp_w = warp / GW, wg = warp % GW;   // stage group, K slice inside a stage
    unsigned long long bar_target = *reinterpret_cast<volatile unsigned long long*>(p.bar_base);
    const int t_cached = __ldg(p.cur_len);         // tokens cached before this step; the new token sits at index t
    int st = 0, pbuf = 0, sq = 0;
    uint32_t ph = 0;

    // rows >= M of the activation operand stay zero for the whole kernel
    for (int i = tid; i < MAXM * p.xpitch / 16; i += CT) sts128(xop + i * 16, make_uint4(0, 0, 0, 0));
    named_bar_sync(1, CT);

    auto stamp = [&](int l, int slot) {
        if (p.prof != nullptr && tid == 0) p.prof[((size_t)cta * p.L + l) * 16 + slot] = global_timer_ns();
    };
    auto grid_barrier = [&](int code) {
        named_bar_sync(1, CT);
        if (tid == 0) {
            __threadfence();
            atomicAdd(p.bar_ctr, 1ull);
            bar_target += (unsigned long long)G;
            unsigned long long v;
            uint32_t spins = 0;
            uint64_t t0 = 0;
            while (true) {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p.bar_ctr) : "memory");
                if (v >= bar_target) break;
                if ((++spins & 0xff) == 0) {
                    const uint64_t now = global_timer_ns();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > CV_WAIT_TIMEOUT_NS) step_fail(p, code);
                }
            }
        } else {
            bar_target += (unsigned long long)G;
        }
        named_bar_sync(1, CT);
    };

    // y[m, n] for the CTA's rows of one matrix, x = the shared-memory operand.  Ring stages alternate between the two
    // warp groups (stage counter `sq` is the producer's), so two stages are in work at any time; every warp keeps its
    // partial [16 x 8] tile and the 16 partials of a tile are summed through shared memory.
    auto consume = [&](const Mat& m, int epi, const bf16* bias, void* out, int64_t ldo) {
        if (p.quiet && tid == 0) *reinterpret_cast<volatile int*>(&flags[1]) = 0;
        for (int r = m.r0; r < m.r1; r += TILE) {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            for (int ks = 0; ks < nks; ++ks) {
                if ((sq & 1) == grp_w) {
                    mbar_wait(&full[st], ph);
                    const uint32_t wa = ring + st * p.stage_bytes + g * p.pitch + (wg * kpw + q * 8) * 2;
                    const uint32_t xa = xop + g * p.xpitch + (ks * p.kstage + wg * kpw + q * 8) * 2;
                    uint4 w0[CPW], w1[CPW], xv[CPW];
#pragma unroll
                    for (int c = 0; c < CPW; ++c) {
                        w0[c] = lds128(wa + c * 64);
                        w1[c] = lds128(wa + 8 * p.pitch + c * 64);
                        xv[c] = lds128(xa + c * 64);
                    }
#pragma unroll
                    for (int c = 0; c < CPW; ++c) {
                        mma_16816(d, w0[c].x, w1[c].x, w0[c].y, w1[c].y, xv[c].x, xv[c].y);
                        mma_16816(d, w0[c].z, w1[c].z, w0[c].w, w1[c].w, xv[c].z, xv[c].w);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[st]);
                } else if (lane == 0) {
                    mbar_arrive(&empty[st]);                // not this group's stage: nothing to read
                }
                ++sq;
                if (++st == p.nst) { st = 0; ph ^= 1; }
            }
            // D fragment: d0,d1 = (row g, cols 2q,2q+1), d2,d3 = (row g+8, ...); row = output column, col = sequence
            float* pw = part + ((pbuf * CW + warp) * TILE) * 8;
            pw[g * 8 + 2 * q] = d[0];
            pw[g * 8 + 2 * q + 1] = d[1];
            pw[(g + 8) * 8 + 2 * q] = d[2];
            pw[(g + 8) * 8 + 2 * q + 1] = d[3];
            named_bar_sync(1, CT);
            if (tid < TILE * 8) {
                const int mi = tid >> 4, nn = tid & 15;
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < CW; ++w) v += part[((pbuf * CW + w) * TILE + nn) * 8 + mi];
                const int n = r + nn;
                if (mi < M && n < m.r1) {
                    if (epi == EPI_F32) {
                        static_cast<float*>(out)[(size_t)mi * ldo + n] = v;
                    } else if (epi == EPI_FIX64) {
                        atomicAdd(static_cast<unsigned long long*>(out) + (size_t)mi * ldo + n,
                                  (unsigned long long)__float2ll_rn(v * FIX_SCALE));
                    } else {
                        if (bias != nullptr) v += __bfloat162float(bias[n]);
                        if (epi == EPI_BF16_GELU) v = gelu_tanh(v);
                        static_cast<bf16*>(out)[(size_t)mi * ldo + n] = __float2bfloat16_rn(v);
                    }
                }
            }
            pbuf ^= 1;
        }
        if (p.quiet && tid == 0) *reinterpret_cast<volatile int*>(&flags[1]) = 1;
    };

    // x operand <- M rows of h bf16 values written by other CTAs
    auto load_x = [&](const bf16* src, int64_t ld, int koff) {
        const int vpr = h / 8;
        for (int i = tid; i < M * vpr; i += CT) {
            const int mi = i / vpr, c = i - mi * vpr;
            sts128(xop + mi * p.xpitch + c * 16, ldcg_u128(src + (size_t)mi * ld + koff + c * 8));
        }
        named_bar_sync(1, CT);
    };

    // Sandwich-LN glue, computed redundantly by every CTA:
    //   v  = res + LN_post(gemm_out / (max|gemm_out| / 8))          (skipped when gemm_out == nullptr)
    //   xn = LN_pre(v / (max|v| / 8))  -> shared-memory operand;   v -> res_out (this CTA's column slice only)
    // res = res_in (fp32, L2) or, when EMB, the embedding wte[ids] + wpe[pos] of the new token.
    // Thread mapping: row = tid / TPR (TPR = 512 / MR threads per sequence), NVT 4-column vectors per thread — every
    // thread works on ONE row, so a row statistic is a 1-value warp reduction + WPR partials through shared memory.
    constexpr int TPR = CT / MR;
    constexpr int WPR = TPR / 32;
    constexpr int NVT = (2560 / 4 + TPR - 1) / TPR;
    struct GlueArgs {
        const bf16 *gemm_out, *g_post, *b_post;
        const long long* acc;          // != nullptr: gemm_out = bf16(acc * 2^-40 + acc_bias) (the 4h->h output)
        const bf16* acc_bias;
        const float* res_in;
        float* res_out;
        const bf16 *g_pre, *b_pre;
        float eps_post, eps_pre;
        int prof_layer;
    };
    auto glue = [&](auto emb_tag, const GlueArgs& a) {
        constexpr bool EMB = decltype(emb_tag)::value;
        const int hv = h >> 2;
        const float inv_h = 1.0f / h;
        const int grow = tid / TPR, gt = tid - grow * TPR;
        const bool row_ok = grow < M;
        float2* red2 = reinterpret_cast<float2*>(red);       // [4][CW] (sum, max)
        auto reduce = [&](float& s, float& mx, int which) {
            s = warp_sum(s);
            mx = warp_max(mx);
            if (lane == 0) red2[which * CW + warp] = make_float2(s, mx);
            named_bar_sync(1, CT);
            float t = 0.f, m2 = 0.f;
#pragma unroll
            for (int w = 0; w < WPR; ++w) t += red2[which * CW + grow * WPR + w].x;
#pragma unroll
            for (int w = 0; w < CW; ++w) m2 = fmaxf(m2, red2[which * CW + w].y);
            s = t;
            mx = m2;
        };
        auto sum4 = [](const float4& x) { return (x.x + x.y) + (x.z + x.w); };
        auto amax4 = [](const float4& x) { return fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))); };
        auto dev4 = [](const float4& x, float mean) {
            const float a0 = x.x - mean, a1 = x.y - mean, a2 = x.z - mean, a3 = x.w - mean;
            return (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        };
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint2* wrow = nullptr;
        const uint2* prow = nullptr;
        if (EMB) {                                   // rows >= M read row 0 (discarded)
            const int rr = row_ok ? grow : 0;
            wrow = reinterpret_cast<const uint2*>(p.wte + (size_t)__ldg(p.ids + rr) * h);
            prow = reinterpret_cast<const uint2*>(p.wpe + (size_t)__ldg(p.pos + rr) * h);
        }
        auto residual = [&](int row, int vi) -> float4 {
            if (!EMB) return __ldcg(reinterpret_cast<const float4*>(a.res_in) + row * hv + vi);
            const uint2 x = __ldg(wrow + vi), y = __ldg(prow + vi);
            return make_float4(bflo(x.x) + bflo(y.x), bfhi(x.x) + bfhi(y.x), bflo(x.y) + bflo(y.y), bfhi(x.y) + bfhi(y.y));
        };

        // Every global load below is UNCONDITIONAL (clamped index, value discarded by a select): loads inside
        // `if (in range)` branches are not hoisted by the compiler, and the first version of this routine paid one
        // serialised L2 round trip per vector (5 x ~0.7 us per pass).
        int vc[NVT];
        bool ok[NVT];
#pragma unroll
        for (int j = 0; j < NVT; ++j) {
            const int vi = gt + TPR * j;
            ok[j] = row_ok && vi < hv;
            vc[j] = ok[j] ? vi : 0;
        }
        const int growc = row_ok ? grow : 0;
        const bool has_post = !EMB && (a.gemm_out != nullptr || a.acc != nullptr);   // layer 0 starts from the embedding
        constexpr bool PRE = MR <= 4;               // small batches: parameters and residual are fetched up front
        uint2 gq[PRE ? NVT : 1], bq[PRE ? NVT : 1];
        if (PRE) {
#pragma unroll
            for (int j = 0; j < NVT; ++j) {
                gq[PRE ? j : 0] = __ldg(reinterpret_cast<const uint2*>(a.g_pre) + vc[j]);
                bq[PRE ? j : 0] = __ldg(reinterpret_cast<const uint2*>(a.b_pre) + vc[j]);
            }
        }
        float4 v[NVT];
        if (has_post) {
            uint2 gp[NVT], bp[NVT];
#pragma unroll
            for (int j = 0; j < NVT; ++j) {
                gp[j] = __ldg(reinterpret_cast<const uint2*>(a.g_post) + vc[j]);
                bp[j] = __ldg(reinterpret_cast<const uint2*>(a.b_post) + vc[j]);
            }
            float s = 0.f, amax = 0.f;
            if (a.acc != nullptr) {
#pragma unroll
                for (int j = 0; j < NVT; ++j) {
                    const longlong2* src = reinterpret_cast<const longlong2*>(a.acc) + (growc * hv + vc[j]) * 2;
                    const longlong2 a01 = __ldcg(src), a23 = __ldcg(src + 1);
                    const uint2 bb = __ldg(reinterpret_cast<const uint2*>(a.acc_bias) + vc[j]);
                    // the 4h->h output as the per-operation path produces it: bf16(sum + bias)
                    const uint32_t lo = pack_bf16x2(__ll2float_rn(a01.x) * FIX_INV + bflo(bb.x),
                                                    __ll2float_rn(a01.y) * FIX_INV + bfhi(bb.x));
                    const uint32_t hi = pack_bf16x2(__ll2float_rn(a23.x) * FIX_INV + bflo(bb.y),
                                                    __ll2float_rn(a23.y) * FIX_INV + bfhi(bb.y));
                    v[j] = ok[j] ? make_float4(bflo(lo), bfhi(lo), bflo(hi), bfhi(hi)) : zero4;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NVT; ++j) {
                    const uint2 u = __ldcg(reinterpret_cast<const uint2*>(a.gemm_out) + growc * hv + vc[j]);
                    v[j] = ok[j] ? make_float4(bflo(u.x), bfhi(u.x), bflo(u.y), bfhi(u.y)) : zero4;
                }
            }
            float4 rs[NVT];
#pragma unroll
            for (int j = 0; j < NVT; ++j) rs[j] = residual(growc, vc[j]);
#pragma unroll
            for (int j = 0; j < NVT; ++j) {
                s += sum4(v[j]);
                amax = fmaxf(amax, amax4(v[j]));
            }
            reduce(s, amax, 0);
            if (a.prof_layer >= 0) stamp(a.prof_layer, 13);
            const float c = amax * 0.125f, mean = s * inv_h;
            float ss = 0.f, dummy = 0.f;
#pragma unroll
            for (int j = 0; j < NVT; ++j)
                if (ok[j]) ss += dev4(v[j], mean);
            reduce(ss, dummy, 1);
            const float rstd = rsqrtf(ss * inv_h + a.eps_post * c * c);
#pragma unroll
            for (int j = 0; j < NVT; ++j) {
                float4 x = v[j];
                x.x = (x.x - mean) * rstd * bflo(gp[j].x) + bflo(bp[j].x) + rs[j].x;
                x.y = (x.y - mean) * rstd * bfhi(gp[j].x) + bfhi(bp[j].x) + rs[j].y;
                x.z = (x.z - mean) * rstd * bflo(gp[j].y) + bflo(bp[j].y) + rs[j].z;
                x.w = (x.w - mean) * rstd * bfhi(gp[j].y) + bfhi(bp[j].y) + rs[j].w;
                v[j] = ok[j] ? x : zero4;
            }
        } else {
#pragma unroll
            for (int j = 0; j < NVT; ++j) {
                const float4 r = residual(growc, vc[j]);
                v[j] = ok[j] ? r : zero4;
            }
        }
        // residual stream out (owner slice), statistics of v
        const int v_lo = (int)(((int64_t)hv * cta) / G), v_hi = (int)(((int64_t)hv * (cta + 1)) / G);
        float s2 = 0.f, amax2 = 0.f;
#pragma unroll
        for (int j = 0; j < NVT; ++j) {
            const int vi = gt + TPR * j;
            if (ok[j] && a.res_out != nullptr && vi >= v_lo && vi < v_hi)
                *(reinterpret_cast<float4*>(a.res_out) + grow * hv + vi) = v[j];
            s2 += sum4(v[j]);
            amax2 = fmaxf(amax2, amax4(v[j]));
        }
        uint2 gq2[PRE ? 1 : NVT], bq2[PRE ? 1 : NVT];
        if (!PRE) {
#pragma unroll
            for (int j = 0; j < NVT; ++j) {
                gq2[PRE ? 0 : j] = __ldg(reinterpret_cast<const uint2*>(a.g_pre) + vc[j]);
                bq2[PRE ? 0 : j] = __ldg(reinterpret_cast<const uint2*>(a.b_pre) + vc[j]);
            }
        }
        reduce(s2, amax2, 2);
        if (a.prof_layer >= 0) stamp(a.prof_layer, 14);
        const float c2 = amax2 * 0.125f, mean2 = s2 * inv_h;
        float ss2 = 0.f, dummy2 = 0.f;
#pragma unroll
        for (int j = 0; j < NVT; ++j)
            if (ok[j]) ss2 += dev4(v[j], mean2);
        reduce(ss2, dummy2, 3);
        if (a.prof_layer >= 0) stamp(a.prof_layer, 15);
        const float rstd2 = rsqrtf(ss2 * inv_h + a.eps_pre * c2 * c2);
#pragma unroll
        for (int j = 0; j < NVT; ++j) {
            const uint2 gg = PRE ? gq[PRE ? j : 0] : gq2[PRE ? 0 : j];
            const uint2 bb = PRE ? bq[PRE ? j : 0] : bq2[PRE ? 0 : j];
            const float4 x = v[j];
            const uint32_t lo = pack_bf16x2((x.x - mean2) * rstd2 * bflo(gg.x) + bflo(bb.x),
                                            (x.y - mean2) * rstd2 * bfhi(gg.x) + bfhi(bb.x));
            const uint32_t hi = pack_bf16x2((x.z - mean2) * rstd2 * bflo(gg.y) + bflo(bb.y),
                                            (x.w - mean2) * rstd2 * bfhi(gg.y) + bfhi(bb.y));
            if (ok[j])
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(xop + grow * p.xpitch + (gt + TPR * j) * 8), "r"(lo),
                             "r"(hi) : "memory");
        }
        named_bar_sync(1, CT);
    };

    // attention of the new token over keys 0..t (standard_attention for sq = 1).  The cached keys reach the CTA through
    // the same ring as the weights: the producer streams [64 keys x 64 dims] K and V tiles of this CTA's share of the
    // flattened (sequence, head, key block) space right after the QKV weights — cached keys do not depend on this
    // step, so they are on their way while QKV is still being multiplied and no DRAM latency is exposed here.
    // A stage is worked on by all 16 warps (4 keys per warp, 8 lanes per key); a warp keeps its online-softmax
    // state across the consecutive blocks of a (sequence, head) pair; at the end of a pair the 16 warp states are
    // merged through shared memory, and pairs that straddle CTAs through L2 partials + an arrival counter
    // (fixed merge order: deterministic).  The new token's K/V come from the QKV output and are appended in place.
    auto attention = [&](int l) {
        bf16* cache_l = p.cache + (size_t)l * p.cache_ls;
        const int grp = lane >> 3, sub = lane & 7;
        const int t = t_cached;
        const int nblk = (t + KVB - 1) / KVB;
        const int pairs = M * p.heads;
        const long long NBt = (long long)pairs * nblk;
        const int f0 = (int)((NBt * cta) / G), f1 = (int)((NBt * (cta + 1)) / G);
        float* sm_acc = part;                                 // [CW][HD] warp states (aliases the tile partials)
        float* sm_ml = red;                                   // [CW][2]
        // CTA that owns flattened block f: the largest c with floor(NBt c / G) <= f  (32-bit: NBt G < 2^31)
        const int nbt = (int)NBt;
        auto owner = [&](int f) { return min(G - 1, ((f + 1) * G - 1) / nbt); };
        float qf[8], m = -INFINITY, lsum = 0.f, acc[8];
        int cur = -1;
        auto start_pair = [&](int bh) {
            const int head = bh % p.heads, batch = bh / p.heads;
            bf16x8_to_float(ldcg_u128(p.qkv + (size_t)batch * 3 * h + head * HD + sub * 8), qf);
            m = -INFINITY;
            lsum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        };
        auto add_key = [&](const uint4& kr, const uint4& vr, bool valid) {
            float kf[8];
            bf16x8_to_float(kr, kf);
            float sdot = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) sdot = fmaf(qf[i], kf[i], sdot);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
            if (valid) {
                const float sc = sdot * p.scale_log2;
                const float mn = fmaxf(m, sc);
                const float alpha = exp2f(m - mn), pr = exp2f(sc - mn);
                float vf[8];
                bf16x8_to_float(vr, vf);
                m = mn;
                lsum = lsum * alpha + pr;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, vf[i], acc[i] * alpha);
            }
        };
        auto finish_pair = [&](int bh) {
            const int head = bh % p.heads, batch = bh / p.heads;
            // contributors of this pair = the distinct owners of its key blocks.  With at least one block per CTA
            // (nbt >= G) they are the consecutive CTAs first..last; otherwise every non-empty CTA owns exactly one block.
            int ncontrib = 1, my_idx = 0;
            bool owns_last = true;
            if (nblk > 0) {
                const int cf = owner(bh * nblk), cl = owner(bh * nblk + nblk - 1);
                owns_last = cl == cta;
                if (nbt >= G) {
                    ncontrib = cl - cf + 1;
                    my_idx = cta - cf;
                } else {
                    ncontrib = nblk;
                    my_idx = f0 - bh * nblk;            // this CTA's single block
                }
            }
            if (owns_last && warp == 0) {                     // the new token: key index t, K/V from the QKV output
                const bf16* qrow = p.qkv + (size_t)batch * 3 * h + head * HD + sub * 8;
                const uint4 knew = ldcg_u128(qrow + h), vnew = ldcg_u128(qrow + 2 * h);
                if (grp == 0 && t < p.max_len) {
                    bf16* kdst = cache_l + (size_t)batch * p.cache_bs + (size_t)t * 2 * h + head * HD + sub * 8;
                    *reinterpret_cast<uint4*>(kdst) = knew;
                    *reinterpret_cast<uint4*>(kdst + h) = vnew;
                }
                add_key(knew, vnew, grp == 0);
            }
            // merge the warp's four key groups (lanes with the same `sub`)
#pragma unroll
            for (int off = 8; off <= 16; off <<= 1) {
                const float mo = __shfl_xor_sync(0xffffffffu, m, off);
                const float lo = __shfl_xor_sync(0xffffffffu, lsum, off);
                const float mn = fmaxf(m, mo);
                const float wa = (m == -INFINITY) ? 0.f : exp2f(m - mn);
                const float wb = (mo == -INFINITY) ? 0.f : exp2f(mo - mn);
                lsum = lsum * wa + lo * wb;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float ao = __shfl_xor_sync(0xffffffffu, acc[i], off);
                    acc[i] = acc[i] * wa + ao * wb;
                }
                m = mn;
            }
            if (grp == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) sm_acc[warp * HD + sub * 8 + i] = acc[i];
                if (sub == 0) { sm_ml[2 * warp] = m; sm_ml[2 * warp + 1] = lsum; }
            }
            named_bar_sync(1, CT);
            if (tid < HD) {
                float Mx = -INFINITY;
#pragma unroll
                for (int w = 0; w < CW; ++w) Mx = fmaxf(Mx, sm_ml[2 * w]);
                float Ls = 0.f, A = 0.f;
#pragma unroll
                for (int w = 0; w < CW; ++w) {
                    const float mw = sm_ml[2 * w];
                    const float wgt = (mw == -INFINITY) ? 0.f : exp2f(mw - Mx);
                    Ls += sm_ml[2 * w + 1] * wgt;
                    A += sm_acc[w * HD + tid] * wgt;
                }
                bf16* o = p.ctx + (size_t)batch * h + head * HD;
                if (ncontrib == 1) {
                    o[tid] = __float2bfloat16_rn(A / Ls);
                } else {
                    float* dst = p.attn_part + ((size_t)bh * KVB_MAX + my_idx) * PART_STRIDE;
                    dst[tid] = A;
                    if (tid == 0) { dst[HD] = Mx; dst[HD + 1] = Ls; }
                    __threadfence();
                    named_bar_sync(2, HD);
                    if (tid == 0) flags[0] = (atomicAdd(p.attn_cnt + bh, 1u) == (unsigned int)(ncontrib - 1)) ? 1 : 0;
                    named_bar_sync(2, HD);
                    if (flags[0]) {                          // last contributor: merge in contributor order
                        __threadfence();
                        const float* src = p.attn_part + (size_t)bh * KVB_MAX * PART_STRIDE;
                        float M2 = -INFINITY;
                        for (int c2 = 0; c2 < ncontrib; ++c2) M2 = fmaxf(M2, __ldcg(src + c2 * PART_STRIDE + HD));
                        float L2 = 0.f, A2 = 0.f;
                        for (int c2 = 0; c2 < ncontrib; ++c2) {
                            const float ms = __ldcg(src + c2 * PART_STRIDE + HD);
                            const float wgt = (ms == -INFINITY) ? 0.f : exp2f(ms - M2);
                            L2 += __ldcg(src + c2 * PART_STRIDE + HD + 1) * wgt;
                            A2 += __ldcg(src + c2 * PART_STRIDE + tid) * wgt;
                        }
                        o[tid] = __float2bfloat16_rn(A2 / L2);
                        if (tid == 0) p.attn_cnt[bh] = 0u;
                    }
                }
            }
            named_bar_sync(1, CT);                            // sm_acc / sm_ml / flags free for the next pair
        };

        if (nblk == 0) {                                      // empty cache: only the new token, pairs round-robin
            for (int bh = cta; bh < pairs; bh += G) {
                start_pair(bh);
                finish_pair(bh);
            }
            return;
        }
        for (int f = f0; f < f1; ++f) {
            const int bh = f / nblk, blk = f - bh * nblk;
            if (bh != cur) {
                if (cur >= 0) finish_pair(cur);
                start_pair(bh);
                cur = bh;
            }
            mbar_wait(&full[st], ph);
            const uint32_t ka = ring + st * p.stage_bytes + (warp * 4 + grp) * 128 + sub * 16;
            const uint4 kr = lds128(ka), vr = lds128(ka + KVB * 128);
            add_key(kr, vr, blk * KVB + warp * 4 + grp < t);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
            ++sq;
            if (++st == p.nst) { st = 0; ph ^= 1; }
        }
        if (cur >= 0) finish_pair(cur);
    };

    // owner slice of a fixed-point accumulator back to zero (its next use is two layers ahead)
    auto zero_acc = [&](long long* acc) {
        const int hv2 = h >> 1;                             // 16-byte (2 x int64) vectors per row
        const int lo = (int)(((int64_t)hv2 * cta) / G), hi = (int)(((int64_t)hv2 * (cta + 1)) / G);
        for (int i = tid; i < MAXM * (hi - lo); i += CT) {
            const int mi = i / (hi - lo), c = lo + (i - mi * (hi - lo));
            reinterpret_cast<longlong2*>(acc)[(size_t)mi * hv2 + c] = make_longlong2(0ll, 0ll);
        }
    };

    // ------------------------------------------------------------------------------------------------
    // the step: 5 phases per layer (each ends in a grid barrier), then final LayerNorm + logits
    //   0: x = x_prev + LN4(mlp_out_prev) (layer 0: embedding), LN1, QKV      1: attention over the K|V cache
    //   2: dense                3: y = x + LN3(attn_out), LN2, h->4h + GELU    4: 4h->h (K quarters) + merge
    // ------------------------------------------------------------------------------------------------
    const int n_it = 5 * p.L + 1;
    const int rg = cta / KG, kg = cta % KG;
    zero_acc(p.fc2_acc);                                    // both accumulators: first use is 4 grid barriers away
    zero_acc(p.fc2_acc + (size_t)MAXM * h);
    for (int it = 0; it < n_it; ++it) {
        const int l = it / 5, phs = it - 5 * l;
        const bool fin = it == n_it - 1;
        const cv_decode_layer& Lw = p.layers[fin ? p.L - 1 : l];
        if (!fin && phs == 0) stamp(l, 0);
        // ---- glue ----
        if (fin || phs == 0 || phs == 3) {
            GlueArgs ga;
            ga.eps_post = p.eps;
            ga.eps_pre = p.eps;
            ga.prof_layer = -1;
            ga.acc = nullptr;
            ga.acc_bias = nullptr;
            if (fin) {
                ga.gemm_out = nullptr; ga.acc = p.fc2_acc + (size_t)((p.L - 1) & 1) * MAXM * h;
                ga.acc_bias = static_cast<const bf16*>(Lw.b_fc2); ga.g_post = static_cast<const bf16*>(Lw.ln4_g); ga.b_post = static_cast<const bf16*>(Lw.ln4_b);
                ga.res_in = p.resid_a; ga.res_out = nullptr;
                ga.g_pre = p.lnf_g; ga.b_pre = p.lnf_b; ga.eps_pre = p.eps_final;
            } else if (phs == 0) {
                const cv_decode_layer& Lp = p.layers[l > 0 ? l - 1 : 0];
                ga.gemm_out = nullptr;
                if (l > 0) {
                    ga.acc = p.fc2_acc + (size_t)((l - 1) & 1) * MAXM * h;
                    ga.acc_bias = static_cast<const bf16*>(Lp.b_fc2);
                }
                ga.g_post = static_cast<const bf16*>(Lp.ln4_g); ga.b_post = static_cast<const bf16*>(Lp.ln4_b);
                ga.res_in = p.resid_a; ga.res_out = p.resid_b;
                ga.g_pre = static_cast<const bf16*>(Lw.ln1_g); ga.b_pre = static_cast<const bf16*>(Lw.ln1_b);
                ga.prof_layer = l > 0 ? l : -1;
            } else {
                ga.gemm_out = p.attn_out; ga.g_post = static_cast<const bf16*>(Lw.ln3_g); ga.b_post = static_cast<const bf16*>(Lw.ln3_b);
                ga.res_in = p.resid_b; ga.res_out = p.resid_a;
                ga.g_pre = static_cast<const bf16*>(Lw.ln2_g); ga.b_pre = static_cast<const bf16*>(Lw.ln2_b);
            }
            if (it == 0) glue(std::true_type{}, ga);
            else glue(std::false_type{}, ga);
            if (!fin) stamp(l, phs == 0 ? 1 : 8);
        }
        // ---- activation operand written by other CTAs ----
        if (!fin && (phs == 2 || phs == 4)) {
            if (phs == 2) {
                if (l > 0) zero_acc(p.fc2_acc + (size_t)((l - 1) & 1) * MAXM * h);   // read by every CTA in phase 0
                load_x(p.ctx, h, 0);
            } else {
                load_x(p.h4, 4 * h, kg * h);
            }
        }
        // ---- the phase's work ----
        if (!fin && phs == 1) {
            attention(l);
        } else {
            Mat m;
            int epi = EPI_BF16;
            const bf16* bias = nullptr;
            void* out = nullptr;
            int64_t ldo = h;
            if (fin) {
                m = make_mat(p.wte, h, p.V, cta, G, 0); epi = EPI_F32; out = p.logits; ldo = p.ldl;
            } else if (phs == 0) {
                m = make_mat(Lw.w_qkv, h, 3 * h, cta, G, 0); bias = static_cast<const bf16*>(Lw.b_qkv); out = p.qkv; ldo = 3 * h;
            } else if (phs == 2) {
                m = make_mat(Lw.w_dense, h, h, cta, G, 0); bias = static_cast<const bf16*>(Lw.b_dense); out = p.attn_out;
            } else if (phs == 3) {
                m = make_mat(Lw.w_fc1, h, 4 * h, cta, G, 0); epi = EPI_BF16_GELU; bias = static_cast<const bf16*>(Lw.b_fc1);
                out = p.h4; ldo = 4 * h;
            } else {
                m = make_mat(Lw.w_fc2, 4 * (int64_t)h, h, rg, G / KG, kg * h); epi = EPI_FIX64;
                out = p.fc2_acc + (size_t)(l & 1) * MAXM * h;
            }
            consume(m, epi, bias, out, ldo);
        }
        if (!fin) {
            const int after_work = phs == 0 ? 2 : (phs == 1 ? 4 : (phs == 2 ? 6 : (phs == 3 ? 9 : 11)));
            stamp(l, after_work);
            grid_barrier(100 * (phs + 1) + l);
            stamp(l, after_work + 1);
        }
    }
    if (cta == 0 && tid == 0) *p.bar_base = bar_target;
}

// workspace layout (bytes); the first WS_ZERO bytes hold counters and must start zeroed
constexpr size_t WS_CTR = 0;        // bar_ctr u64, bar_base u64, err int
constexpr size_t WS_FC2CNT = 64;    // u32[64]
constexpr size_t WS_ATTNCNT = 320;  // u32[MAXM * heads]
constexpr size_t WS_DATA = 8192;
inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

struct WsLayout {
    size_t qkv, ctx, attn_out, h4, mlp_out, resid_a, resid_b, fc2_acc, attn_part, total;
};
WsLayout ws_layout(int h, int heads) {
    WsLayout w;
    size_t o = WS_DATA;
    w.qkv = o; o += al256((size_t)MAXM * 3 * h * 2);
    w.ctx = o; o += al256((size_t)MAXM * h * 2);
    w.attn_out = o; o += al256((size_t)MAXM * h * 2);
    w.h4 = o; o += al256((size_t)MAXM * 4 * h * 2);
    w.mlp_out = o; o += al256((size_t)MAXM * h * 2);
    w.resid_a = o; o += al256((size_t)MAXM * h * 4);
    w.resid_b = o; o += al256((size_t)MAXM * h * 4);
    w.fc2_acc = o; o += al256((size_t)2 * MAXM * h * 8);
    w.attn_part = o; o += al256((size_t)MAXM * heads * KVB_MAX * PART_STRIDE * 4);
    w.total = o;
    return w;
}

int step_grid() {
    int g = cvh::num_sms();
    return g - g % KG;
}

}  // namespace

extern "C" int64_t cv_decode_step_workspace_bytes(int hidden, int heads) {
    if (hidden <= 0 || heads <= 0) return -1;
    return (int64_t)ws_layout(hidden, heads).total;
}

extern "C" int cv_decode_step(const cv_decode_step_args* a, void* stream) {
    CV_REQUIRE(a != nullptr && a->layers && a->wte && a->wpe && a->lnf_g && a->lnf_b && a->ids && a->pos &&
                   a->cur_len && a->cache && a->logits && a->workspace,
               "null pointer");
    const int h = a->hidden, heads = a->heads, M = a->batch;
    CV_REQUIRE(M >= 1 && M <= MAXM, "cv_decode_step handles 1 <= batch <= 8 sequences");
    CV_REQUIRE(h > 0 && h % 256 == 0 && h <= 2560, "hidden must be a multiple of 256 and <= 2560");
    CV_REQUIRE(heads > 0 && heads * HD == h && heads <= (1024 / MAXM), "hidden must be heads * 64");
    CV_REQUIRE(a->num_layers >= 1 && a->vocab >= 1 && a->max_len >= 1 && a->ld_logits >= a->vocab, "bad sizes");
    CV_REQUIRE(a->max_len <= KVB * KVB_MAX, "max_len must be <= 4096");
    CV_REQUIRE(a->cache_batch_stride == (int64_t)a->max_len * 2 * h &&
                   a->cache_layer_stride == (int64_t)M * a->max_len * 2 * h &&
                   (reinterpret_cast<uintptr_t>(a->cache) & 15) == 0,
               "the K|V cache must be a contiguous [layers, batch, max_len, 2*hidden] bf16 tensor");
    CV_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "workspace must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = step_grid();
    CV_REQUIRE(grid >= KG && grid / KG <= 64, "unsupported SM count");

    Params p;
    p.layers = a->layers;
    p.L = a->num_layers; p.h = h; p.heads = heads; p.V = a->vocab; p.M = M; p.max_len = a->max_len;
    p.eps = a->eps; p.eps_final = a->eps_final;
    p.wte = static_cast<const bf16*>(a->wte); p.wpe = static_cast<const bf16*>(a->wpe);
    p.lnf_g = static_cast<const bf16*>(a->lnf_g); p.lnf_b = static_cast<const bf16*>(a->lnf_b);
    p.ids = a->ids; p.pos = a->pos; p.cur_len = a->cur_len;
    p.cache = static_cast<bf16*>(a->cache);
    p.cache_ls = a->cache_layer_stride; p.cache_bs = a->cache_batch_stride;
    p.logits = a->logits; p.ldl = a->ld_logits;
    p.prof = static_cast<unsigned long long*>(a->prof);
    char* ws = static_cast<char*>(a->workspace);
    const WsLayout w = ws_layout(h, heads);
    p.bar_ctr = reinterpret_cast<unsigned long long*>(ws + WS_CTR);
    p.bar_base = p.bar_ctr + 1;
    p.err = reinterpret_cast<int*>(ws + WS_CTR + 16);
    p.attn_cnt = reinterpret_cast<unsigned int*>(ws + WS_ATTNCNT);
    p.qkv = reinterpret_cast<bf16*>(ws + w.qkv); p.ctx = reinterpret_cast<bf16*>(ws + w.ctx);
    p.attn_out = reinterpret_cast<bf16*>(ws + w.attn_out); p.h4 = reinterpret_cast<bf16*>(ws + w.h4);
    p.mlp_out = reinterpret_cast<bf16*>(ws + w.mlp_out);
    p.resid_a = reinterpret_cast<float*>(ws + w.resid_a); p.resid_b = reinterpret_cast<float*>(ws + w.resid_b);
    p.fc2_acc = reinterpret_cast<long long*>(ws + w.fc2_acc); p.attn_part = reinterpret_cast<float*>(ws + w.attn_part);

    // stage = 16 weight rows x kstage columns; kstage = the largest multiple of 256 dividing h that is <= 1280
    int kstage = 256;
    for (int k = 256; k <= 1280; k += 256)
        if (h % k == 0) kstage = k;
    p.kstage = kstage;
    p.pitch = kstage * 2 + 64;                       // 64 mod 128: conflict-free 16-byte fragment loads
    p.xpitch = h * 2 + 64;
    p.stage_bytes = TILE * p.pitch;
    const int fixed = SM_XOP + ((MAXM * p.xpitch + 127) / 128) * 128;
    int max_smem = 0, dev = 0;
    CV_CUDA(cudaGetDevice(&dev));
    CV_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int nst = (max_smem - fixed - 1024) / p.stage_bytes;
    if (nst > MAXST) nst = MAXST;
    if (const char* e = getenv("COGVIEW_B200_STEP_NST")) {
        const int cap = atoi(e);
        if (cap >= 2 && cap < nst) nst = cap;
    }
    {
        const char* e = getenv("COGVIEW_B200_STEP_QUIET");
        p.quiet = (e && e[0] == '1') ? 1 : 0;
        // L2 prefetch distance in ring stages beyond the ring (~41 KB each, x SMs): 10 -> ~60 MB of the 126 MB L2
        const char* f = getenv("COGVIEW_B200_STEP_PF");
        p.pf_stages = f ? atoi(f) : 10;
        if (p.pf_stages < 0) p.pf_stages = 0;
        if (p.pf_stages > 64) p.pf_stages = 64;
    }
    CV_REQUIRE(nst >= 2, "not enough shared memory for the weight ring");
    p.nst = nst;
    const size_t smem_bytes = (size_t)fixed + (size_t)nst * p.stage_bytes;
    int S = (grid * CW) / (M * heads);
    S = S < 1 ? 1 : (S > 16 ? 16 : S);
    p.S = S;
    p.scale_log2 = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;

    // K|V cache as a 4-D tensor: [2h | max_len | batch | layer], boxes of [64 dims x KVB keys] (one head's K or V)
    alignas(64) CUtensorMap tmKV;
    {
        const uint64_t dims[4] = {(uint64_t)2 * h, (uint64_t)a->max_len, (uint64_t)M, (uint64_t)a->num_layers};
        const uint64_t str[3] = {(uint64_t)2 * h * 2, (uint64_t)a->cache_batch_stride * 2, (uint64_t)a->cache_layer_stride * 2};
        const uint32_t box[4] = {HD, KVB, 1, 1};
        int rc = cvh::encode_tmap(&tmKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->cache, dims, str, box, nullptr,
                                  cvh::Swizzle::None);
        if (rc) return rc;
    }
    typedef void (*KernelFn)(const Params, const CUtensorMap);
    KernelFn fn = nullptr;
    const int cpw = kstage / 256;
#define DS_PICK(MR_)                                                    \
    switch (cpw) {                                                      \
        case 1: fn = decode_step_kernel<MR_, 1>; break;                 \
        case 2: fn = decode_step_kernel<MR_, 2>; break;                 \
        case 3: fn = decode_step_kernel<MR_, 3>; break;                 \
        case 4: fn = decode_step_kernel<MR_, 4>; break;                 \
        default: fn = decode_step_kernel<MR_, 5>; break;                \
    }
    if (M <= 4) { DS_PICK(4) } else { DS_PICK(8) }
#undef DS_PICK
    CV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    static int coop = -1;
    if (coop < 0) {
        const char* e = getenv("COGVIEW_B200_COOP");
        int sup = 0;
        cudaDeviceGetAttribute(&sup, cudaDevAttrCooperativeLaunch, dev);
        coop = (sup && !(e && e[0] == '0')) ? 1 : 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = coop ? 1 : 0;
    CV_CUDA(cudaLaunchKernelEx(&cfg, fn, p, tmKV));
    cvh::count_launches(1);
    return 0;
}
