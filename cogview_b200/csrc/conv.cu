// VQ-VAE convolutions as im2col-free implicit GEMMs on tcgen05 (NHWC bf16 activations).
//
//   cv_conv2d_k4s2           nn.Conv2d(Cin, Cout, 4, stride=2, padding=1)          encoder, /root/reference/vqvae/vqvae_zc.py:121-129
//   cv_conv_transpose2d_k4s2 nn.ConvTranspose2d(Cin, Cout, 4, stride=2, padding=1) decoder, vqvae/vqvae_zc.py:172-191
//
// GEMM view: rows = output pixels (128 per tile), columns = output channels, K = taps x Cin.  The A tile of one
// tap is one TMA box of the NHWC input — [64 channels x TW x TH x NB] with traversal stride 2 in W and H for the
// strided convolution (elementStrides), out-of-bounds coordinates zero-filled by TMA (that IS the padding) — so
// no im2col buffer is ever materialised.  Weights are pre-packed [tap][Cout][Cin] (K-major B tiles).
// The transposed convolution is computed as its 4 sub-pixel phases (each a 2x2-tap stride-1 convolution on the
// input grid); a phase's output tile is scattered to (2a+py, 2b+px) with one 5-D TMA store.
// Pipeline / warp roles / TMEM double-buffering are those of gemm.cu; epilogue = bias (+ReLU) -> bf16 -> TMA store.
#include "common.cuh"
#include "host.h"
#include "../../include/cogview_b200.h"

namespace {
using namespace cv;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NUM_THREADS = 192;
constexpr int EPI_BYTES = 128 * 128;

template <int BN>
struct Cfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * EPI_BYTES + 1024 + 256;
};

struct ConvParams {
    int Cin, Cout;
    int H, W;            // tile grid: output dims for the strided conv, input dims for a transposed-conv phase
    int TH, NB;          // tile = NB images x TH rows x W columns = 128 pixels
    int tiles_per_image; // (H*W)/128 when >= 1 (then NB == 1)
    int num_m_tiles, num_n_blocks, kc_blocks, ntaps;
    int py, px;          // transposed conv: output phase
    const __nv_bfloat16* bias;
    int relu;
};

// MODE 1: conv k4 s2 p1.  MODE 2: one phase of convT k4 s2 p1.
template <int BN, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const ConvParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_buf = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + 2 * EPI_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tmem_full = bars + 2 * C::STAGES;
    uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_blocks;
    const int num_k_blocks = p.ntaps * p.kc_blocks;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmC);
        for (int i = 0; i < C::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        fence_barrier_init();
    }
    if (warp_idx == 1) tmem_alloc<2 * BN>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto tile_origin = [&](int mt, int& b0, int& y0) {
        if (p.tiles_per_image >= 1) { b0 = mt / p.tiles_per_image; y0 = (mt % p.tiles_per_image) * p.TH; }
        else { b0 = mt * p.NB; y0 = 0; }
    };

    if (warp_idx == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = tile % p.num_m_tiles;
                const int n0 = (tile / p.num_m_tiles) * BN;
                int b0, y0;
                tile_origin(mt, b0, y0);
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    const int tap = kb / p.kc_blocks, kc = kb - tap * p.kc_blocks;
                    int ax, ay, wtap;
                    if (MODE == 1) {
                        const int ky = tap >> 2, kx = tap & 3;
                        ax = kx - 1;                     // 2*0 - 1 + kx (tiles span the full output width)
                        ay = 2 * y0 - 1 + ky;
                        wtap = tap;
                    } else {
                        const int ty = tap >> 1, tx = tap & 1;
                        // output parity 0 uses kernel rows {1,3} at input offsets {0,-1}; parity 1 uses {0,2} at {+1,0}
                        const int ky = p.py == 0 ? (ty == 0 ? 1 : 3) : (ty == 0 ? 0 : 2);
                        const int kx = p.px == 0 ? (tx == 0 ? 1 : 3) : (tx == 0 ? 0 : 2);
                        const int dy = p.py == 0 ? (ty == 0 ? 0 : -1) : (ty == 0 ? 1 : 0);
                        const int dx = p.px == 0 ? (tx == 0 ? 0 : -1) : (tx == 0 ? 1 : 0);
                        ax = dx;
                        ay = y0 + dy;
                        wtap = ky * 4 + kx;
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * C::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    tma_load_4d(sA, &tmA, &full_bar[stage], kc * BK, ax, ay, b0);
                    tma_load_2d(sA + C::A_BYTES, &tmB, &full_bar[stage], kc * BK, wtap * p.Cout + n0);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_f16(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 0, 1024),
                                 make_smem_desc_sw128(b_addr + k * 32, 0, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[as]);
            }
        }
    } else {
        const int q = warp_idx & 3;
        const int row = q * 32 + lane;
        const int epi_tid = threadIdx.x - 64;
        constexpr int NCHUNK = BN / 64;
        int it = 0;
        uint32_t buf_sel = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int mt = tile % p.num_m_tiles;
            const int n0 = (tile / p.num_m_tiles) * BN;
            int b0, y0;
            tile_origin(mt, b0, y0);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c) {
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * BN + c * 64;
                uint32_t r[64];
                {
                    uint32_t (&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
                    uint32_t (&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
                    tmem_ld_x32(taddr, r0);
                    tmem_ld_x32(taddr + 32, r1);
                }
                tmem_ld_wait();
                if (c == NCHUNK - 1) { tc_fence_before(); mbar_arrive(&tmem_empty[as]); }
                const int ncol0 = n0 + c * 64;
                float v[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    float x = __uint_as_float(r[j]);
                    if (p.bias != nullptr) x += __bfloat162float(p.bias[ncol0 + j]);
                    v[j] = p.relu ? fmaxf(x, 0.f) : x;
                }
                uint8_t* buf = epi_buf + (buf_sel & 1) * EPI_BYTES;
                if (epi_tid == 0) tma_store_wait_read<1>();
                named_bar_sync(1, 128);
                uint8_t* rowp = buf + row * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint4 o;
                    o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                    o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                    o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                    o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                    *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) = o;
                }
                fence_proxy_async_smem();
                named_bar_sync(2, 128);
                if (epi_tid == 0) {
                    if (MODE == 1) tma_store_2d(&tmC, buf, ncol0, mt * BM);
                    else tma_store_5d(&tmC, buf, ncol0, p.px, 0, p.py, b0 * p.H + y0);
                    tma_store_commit();
                }
                ++buf_sel;
            }
        }
        if (epi_tid == 0) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc<2 * BN>(tmem_base);
    }
}

bool pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

template <int BN, int MODE>
int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const ConvParams& p,
                cudaStream_t s) {
    auto kern = conv_kernel<BN, MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        CV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
        attr_set = true;
    }
    const int tiles = p.num_m_tiles * p.num_n_blocks;
    const int grid = tiles < cvh::num_sms() ? tiles : cvh::num_sms();
    kern<<<grid, NUM_THREADS, Cfg<BN>::SMEM_BYTES, s>>>(tmA, tmB, tmC, p);
    CV_LAUNCH_CHECK();
    return 0;
}

// fills the tile decomposition for a [B, H, W] pixel grid; returns false if it cannot be tiled by 128
bool plan_tiles(int B, int H, int W, ConvParams& p) {
    if (!pow2(W) || !pow2(H) || W > 128) return false;
    p.H = H; p.W = W;
    if (H * W >= 128) {
        p.NB = 1; p.TH = 128 / W; p.tiles_per_image = (H * W) / 128;
        p.num_m_tiles = B * p.tiles_per_image;
    } else {
        p.NB = 128 / (H * W); p.TH = H; p.tiles_per_image = 0;
        if (B % p.NB != 0) return false;
        p.num_m_tiles = B / p.NB;
    }
    return true;
}

}  // namespace

extern "C" int cv_conv2d_k4s2(const void* x, const void* w_packed, const void* bias, void* y, int B, int IH, int IW,
                              int Cin, int Cout, int relu, void* stream) {
    CV_REQUIRE(x && w_packed && y, "null pointer");
    CV_REQUIRE(Cin % 64 == 0 && Cout % 128 == 0, "Cin must be a multiple of 64 and Cout of 128");
    CV_REQUIRE(IH % 2 == 0 && IW % 2 == 0, "input height/width must be even");
    const int OH = IH / 2, OW = IW / 2;
    ConvParams p = {};
    CV_REQUIRE(plan_tiles(B, OH, OW, p), "output H, W must be powers of two, W <= 128, and 128 pixels must tile the batch");
    p.Cin = Cin; p.Cout = Cout; p.kc_blocks = Cin / 64; p.ntaps = 16; p.bias = static_cast<const __nv_bfloat16*>(bias);
    p.relu = relu;
    const int BN = (Cout % 256 == 0) ? 256 : 128;
    p.num_n_blocks = Cout / BN;
    alignas(64) CUtensorMap tmA, tmB, tmC;
    {   // input NHWC as [C, W, H, B], traversal stride 2 in W and H
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)IW, (uint64_t)IH, (uint64_t)B};
        uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)IW * Cin * 2, (uint64_t)IH * IW * Cin * 2};
        uint32_t box[4] = {64, (uint32_t)(2 * OW), (uint32_t)(2 * p.TH), (uint32_t)p.NB};
        uint32_t es[4] = {1, 2, 2, 1};
        int rc = cvh::encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box, es, cvh::Swizzle::B128);
        if (rc) return rc;
    }
    int rc = cvh::encode_tmap_2d_bf16(&tmB, w_packed, (uint64_t)16 * Cout, Cin, Cin, BN, 64);
    if (rc) return rc;
    rc = cvh::encode_tmap_2d_bf16(&tmC, y, (uint64_t)B * OH * OW, Cout, Cout, BM, 64);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return BN == 256 ? launch_conv<256, 1>(tmA, tmB, tmC, p, s) : launch_conv<128, 1>(tmA, tmB, tmC, p, s);
}

extern "C" int cv_conv_transpose2d_k4s2(const void* x, const void* w_packed, const void* bias, void* y, int B, int IH,
                                        int IW, int Cin, int Cout, int relu, void* stream) {
    CV_REQUIRE(x && w_packed && y, "null pointer");
    CV_REQUIRE(Cin % 64 == 0 && Cout % 128 == 0, "Cin must be a multiple of 64 and Cout of 128");
    ConvParams p = {};
    CV_REQUIRE(plan_tiles(B, IH, IW, p), "input H, W must be powers of two, W <= 128, and 128 pixels must tile the batch");
    p.Cin = Cin; p.Cout = Cout; p.kc_blocks = Cin / 64; p.ntaps = 4; p.bias = static_cast<const __nv_bfloat16*>(bias);
    p.relu = relu;
    const int BN = (Cout % 256 == 0) ? 256 : 128;
    p.num_n_blocks = Cout / BN;
    alignas(64) CUtensorMap tmA, tmB, tmC;
    {   // input NHWC as [C, W, H, B], unit strides; halo taps fall outside and are zero-filled
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)IW, (uint64_t)IH, (uint64_t)B};
        uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)IW * Cin * 2, (uint64_t)IH * IW * Cin * 2};
        uint32_t box[4] = {64, (uint32_t)IW, (uint32_t)p.TH, (uint32_t)p.NB};
        int rc = cvh::encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box, nullptr,
                                  cvh::Swizzle::B128);
        if (rc) return rc;
    }
    int rc = cvh::encode_tmap_2d_bf16(&tmB, w_packed, (uint64_t)16 * Cout, Cin, Cin, BN, 64);
    if (rc) return rc;
    {   // output [B, 2IH, 2IW, Cout] viewed as [C, px, b, py, (B*IH)]
        const uint64_t OW = 2 * (uint64_t)IW;
        uint64_t dims[5] = {(uint64_t)Cout, 2, (uint64_t)IW, 2, (uint64_t)B * IH};
        uint64_t str[4] = {(uint64_t)Cout * 2, (uint64_t)2 * Cout * 2, OW * Cout * 2, 2 * OW * Cout * 2};
        uint32_t box[5] = {64, 1, (uint32_t)IW, 1, (uint32_t)(p.TH * p.NB)};
        rc = cvh::encode_tmap(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, y, dims, str, box, nullptr, cvh::Swizzle::B128);
        if (rc) return rc;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int ph = 0; ph < 4; ++ph) {
        p.py = ph >> 1; p.px = ph & 1;
        rc = BN == 256 ? launch_conv<256, 2>(tmA, tmB, tmC, p, s) : launch_conv<128, 2>(tmA, tmB, tmC, p, s);
        if (rc) return rc;
    }
    return 0;
}
