// Output side of the voxel path: reconstruction (R1-R2), tri-planar vote (R3), confusion counts (R4).
//
//  recon          predicted 2-D masks (slice orientation, uint8, pixel > 0 == lesion) -> [Z][Y][X] volumes.
//                 Reference: scripts/reconstruir_volumen.py:146-148 (binarise), :179-186 (insertar_corte),
//                 :199-213 (zero volume + loop).  Formulated as a GATHER through an inverse slice map so
//                 every output byte is written exactly once (no memset + scatter): each CTA transposes a
//                 64x64 byte tile through shared memory, reads run along the slices' fastest axis and
//                 writes run along x.
//  consensus_eval (ax + co + sa >= umbral) fused with the 4x4 confusion counts of the three planes and the
//                 consensus against the ground truth.  Reference: scripts/generar_consenso.py:106-109 and
//                 the boolean sums of utils/utils.py:455-495.  SIMD-within-a-register byte predicates,
//                 popc, warp shuffles, one int64 atomic per counter per CTA.
#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

namespace {

// ------------------------------------------------------------------------------------ recon
__global__ void fill_i32_kernel(int32_t* p, size_t n, int32_t val) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = val;
}

__global__ void slot_map_kernel(const int32_t* __restrict__ vol_of_slice, const int32_t* __restrict__ idx_of_slice,
                                int nslices, int nvol, int n_plane, int32_t* __restrict__ slot_of) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslices) return;
    int v = vol_of_slice[s], i = idx_of_slice[s];
    if (v < 0 || v >= nvol || i < 0 || i >= n_plane) return;
    atomicMax(&slot_of[(size_t)v * n_plane + i], s);     // duplicate index: the later slice wins
}

constexpr int kTile = 64;
constexpr int kTilePitch = 68;     // 17 words: conflict-free column reads

struct ReconArgs {
    const uint8_t* slices;
    size_t slice_pitch;
    const int32_t* slot_of;
    uint8_t* vol_u8;
    float* vol_f32;
    int X, Y, Z, plano;
};

// grid (tiles_x * tiles_w, T, nvol); block 256.
//   axial   : T = Z (t = z), w = y   src = slice(z)[x * Y + y]
//   coronal : T = Y (t = y), w = z   src = slice(y)[x * Z + z]
//   sagital : T = Y (t = y), w = z   src = slice(x)[y * Z + z]
__global__ void __launch_bounds__(256) recon_gather_kernel(const ReconArgs a) {
    __shared__ uint8_t tile[kTile][kTilePitch];
    const int X = a.X, Y = a.Y, Z = a.Z;
    const int W = a.plano == MSL_AXIAL ? Y : Z;
    const int n_plane = a.plano == MSL_AXIAL ? Z : (a.plano == MSL_CORONAL ? Y : X);
    const int tiles_x = (X + kTile - 1) / kTile;
    const int tx = blockIdx.x % tiles_x, tw = blockIdx.x / tiles_x;
    const int x0 = tx * kTile, w0 = tw * kTile;
    const int t = blockIdx.y, v = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t* slot_v = a.slot_of + (size_t)v * n_plane;

    int slot_t = a.plano == MSL_SAGITAL ? 0 : slot_v[t];
    for (int xr = warp; xr < kTile; xr += 8) {
        const int x = x0 + xr;
        int slot = -1;
        size_t base = 0;
        if (x < X) {
            if (a.plano == MSL_SAGITAL) { slot = slot_v[x]; base = (size_t)t * W; }
            else { slot = slot_t; base = (size_t)x * W; }
        }
        const uint8_t* src = a.slices + (slot < 0 ? 0 : (size_t)slot * a.slice_pitch) + base;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int w = w0 + h * 32 + lane;
            uint8_t q = 0;
            if (slot >= 0 && w < W) q = __ldg(src + w) > 0 ? 1 : 0;
            tile[xr][h * 32 + lane] = q;
        }
    }
    __syncthreads();
    for (int wr = warp; wr < kTile; wr += 8) {
        const int w = w0 + wr;
        if (w >= W) break;
        const int z = a.plano == MSL_AXIAL ? t : w;
        const int y = a.plano == MSL_AXIAL ? w : t;
        const size_t off = (((size_t)v * Z + z) * Y + y) * X;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int xr = h * 32 + lane, x = x0 + xr;
            if (x < X) {
                uint8_t q = tile[xr][wr];
                if (a.vol_u8) a.vol_u8[off + x] = q;
                if (a.vol_f32) a.vol_f32[off + x] = (float)q;
            }
        }
    }
}

// ------------------------------------------------------------------------------------ vote + counts
// bit 7 of each byte set iff that byte == 0 (exact, no cross-byte borrow)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v) {
    return ~(((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
}
__device__ __forceinline__ uint32_t one_bytes(uint32_t v) { return zero_bytes(v ^ 0x01010101u); }

// (a + b + c >= umbral) per byte -> 0/1 bytes
__device__ __forceinline__ uint32_t vote4(uint32_t a, uint32_t b, uint32_t c, int umbral) {
    if (((a | b | c) & 0xfefefefeu) == 0) {       // all bytes binary: sums <= 3, no carries
        uint32_t s = a + b + c;
        if (umbral == 2) return (s >> 1) & 0x01010101u;
        if (umbral == 3) return (s & (s >> 1)) & 0x01010101u;
        if (umbral == 1) return (s | (s >> 1)) & 0x01010101u;
        return umbral <= 0 ? 0x01010101u : 0u;
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int s = (int)((a >> (8 * k)) & 0xff) + (int)((b >> (8 * k)) & 0xff) + (int)((c >> (8 * k)) & 0xff);
        r |= (uint32_t)(s >= umbral) << (8 * k);
    }
    return r;
}

struct Counts4 { int tp, fp, fn, tn; };
__device__ __forceinline__ void count4(Counts4& c, uint32_t g1, uint32_t g0, uint32_t p) {
    uint32_t p1 = one_bytes(p), p0 = zero_bytes(p);
    c.tp += __popc(g1 & p1); c.fp += __popc(g0 & p1);
    c.fn += __popc(g1 & p0); c.tn += __popc(g0 & p0);
}

constexpr int kCntThreads = 256;

template <int NPLANE>
__device__ __forceinline__ void flush_counts(Counts4 (&c)[NPLANE], long long* dst) {
    __shared__ int s_red[kCntThreads / 32][NPLANE * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int pl = 0; pl < NPLANE; ++pl) {
        int tp = warp_sum(c[pl].tp), fp = warp_sum(c[pl].fp), fn = warp_sum(c[pl].fn), tn = warp_sum(c[pl].tn);
        if (lane == 0) { s_red[warp][pl * 4] = tp; s_red[warp][pl * 4 + 1] = fp; s_red[warp][pl * 4 + 2] = fn; s_red[warp][pl * 4 + 3] = tn; }
    }
    __syncthreads();
    if (threadIdx.x < NPLANE * 4) {
        long long tot = 0;
        for (int w = 0; w < kCntThreads / 32; ++w) tot += s_red[w][threadIdx.x];
        if (tot) atomicAdd(reinterpret_cast<unsigned long long*>(dst) + threadIdx.x, (unsigned long long)tot);
    }
}

struct VoteArgs {
    const uint8_t *ax, *co, *sa, *gt;
    uint8_t* consenso;
    long long* counts;
    size_t nvox;
    int umbral;
};

// grid (chunks, nvol).  VEC: 8-byte granules (nvox % 8 == 0 and 8-byte aligned bases) or bytes.
template <bool VEC>
__global__ void __launch_bounds__(kCntThreads) consensus_eval_kernel(const VoteArgs a) {
    const int v = blockIdx.y;
    const size_t off = (size_t)v * a.nvox;
    Counts4 c[4] = {};
    const bool cnt = a.gt != nullptr;
    if (VEC) {
        const size_t ng = a.nvox / 8;
        const uint2* ax = reinterpret_cast<const uint2*>(a.ax + off);
        const uint2* co = reinterpret_cast<const uint2*>(a.co + off);
        const uint2* sa = reinterpret_cast<const uint2*>(a.sa + off);
        const uint2* gt = cnt ? reinterpret_cast<const uint2*>(a.gt + off) : nullptr;
        uint2* out = a.consenso ? reinterpret_cast<uint2*>(a.consenso + off) : nullptr;
        for (size_t g = (size_t)blockIdx.x * kCntThreads + threadIdx.x; g < ng; g += (size_t)gridDim.x * kCntThreads) {
            uint2 wa = __ldg(ax + g), wc = __ldg(co + g), ws = __ldg(sa + g);
            uint2 wg = cnt ? __ldg(gt + g) : make_uint2(0, 0);
            uint2 r;
            r.x = vote4(wa.x, wc.x, ws.x, a.umbral);
            r.y = vote4(wa.y, wc.y, ws.y, a.umbral);
            if (out) out[g] = r;
            if (cnt) {
                uint32_t g1 = one_bytes(wg.x), g0 = zero_bytes(wg.x);
                count4(c[0], g1, g0, wa.x); count4(c[1], g1, g0, wc.x); count4(c[2], g1, g0, ws.x); count4(c[3], g1, g0, r.x);
                g1 = one_bytes(wg.y); g0 = zero_bytes(wg.y);
                count4(c[0], g1, g0, wa.y); count4(c[1], g1, g0, wc.y); count4(c[2], g1, g0, ws.y); count4(c[3], g1, g0, r.y);
            }
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * kCntThreads + threadIdx.x; i < a.nvox; i += (size_t)gridDim.x * kCntThreads) {
            uint32_t pa = a.ax[off + i], pc = a.co[off + i], ps = a.sa[off + i];
            uint32_t r = ((int)(pa + pc + ps) >= a.umbral) ? 1u : 0u;
            if (a.consenso) a.consenso[off + i] = (uint8_t)r;
            if (cnt) {
                // pad the upper three bytes with 2 so they match neither the ==0 nor the ==1 predicate
                uint32_t g = a.gt[off + i] | 0x02020200u;
                uint32_t g1 = one_bytes(g), g0 = zero_bytes(g);
                count4(c[0], g1, g0, pa | 0x02020200u); count4(c[1], g1, g0, pc | 0x02020200u);
                count4(c[2], g1, g0, ps | 0x02020200u); count4(c[3], g1, g0, r | 0x02020200u);
            }
        }
    }
    if (cnt) flush_counts<4>(c, a.counts + (size_t)v * 16);
}

template <bool VEC>
__global__ void __launch_bounds__(kCntThreads) confusion_counts_kernel(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pred,
                                                                       size_t nvox, long long* __restrict__ counts) {
    const int v = blockIdx.y;
    const size_t off = (size_t)v * nvox;
    Counts4 c[1] = {};
    if (VEC) {
        const size_t ng = nvox / 8;
        const uint2* g8 = reinterpret_cast<const uint2*>(gt + off);
        const uint2* p8 = reinterpret_cast<const uint2*>(pred + off);
        for (size_t g = (size_t)blockIdx.x * kCntThreads + threadIdx.x; g < ng; g += (size_t)gridDim.x * kCntThreads) {
            uint2 wg = __ldg(g8 + g), wp = __ldg(p8 + g);
            count4(c[0], one_bytes(wg.x), zero_bytes(wg.x), wp.x);
            count4(c[0], one_bytes(wg.y), zero_bytes(wg.y), wp.y);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * kCntThreads + threadIdx.x; i < nvox; i += (size_t)gridDim.x * kCntThreads) {
            uint32_t g = gt[off + i] | 0x02020200u, p = pred[off + i] | 0x02020200u;
            count4(c[0], one_bytes(g), zero_bytes(g), p);
        }
    }
    flush_counts<1>(c, counts + (size_t)v * 4);
}

inline bool aligned8(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

inline int chunks_for(size_t nvox, int nvol) {
    // ~16 granules of 8 bytes per thread; at least a couple of waves over 148 SMs for small batches
    size_t per_cta = (size_t)kCntThreads * 8 * 16;
    size_t c = (nvox + per_cta - 1) / per_cta;
    if (c < 1) c = 1;
    if (c > 4096) c = 4096;
    (void)nvol;
    return (int)c;
}

}  // namespace

int launch_recon(const uint8_t* slices, size_t slice_pitch, const int32_t* vol_of_slice, const int32_t* idx_of_slice,
                 int nslices, int plano, int nvol, int X, int Y, int Z, uint8_t* vol_u8, float* vol_f32,
                 int32_t* slot_of, cudaStream_t stream) {
    const int n_plane = plano == MSL_AXIAL ? Z : (plano == MSL_CORONAL ? Y : X);
    const size_t nmap = (size_t)nvol * n_plane;
    { ProfScope prof(K_RECON_FILL, stream);
    fill_i32_kernel<<<(unsigned)((nmap + 255) / 256), 256, 0, stream>>>(slot_of, nmap, -1);
    }
    MSL_LAUNCH_CHECK("fill_i32_kernel");
    if (nslices > 0) {
        ProfScope prof(K_RECON_SLOT_MAP, stream);
        slot_map_kernel<<<(nslices + 255) / 256, 256, 0, stream>>>(vol_of_slice, idx_of_slice, nslices, nvol, n_plane, slot_of);
        MSL_LAUNCH_CHECK("slot_map_kernel");
    }
    ReconArgs a;
    a.slices = slices; a.slice_pitch = slice_pitch; a.slot_of = slot_of; a.vol_u8 = vol_u8; a.vol_f32 = vol_f32;
    a.X = X; a.Y = Y; a.Z = Z; a.plano = plano;
    const int W = plano == MSL_AXIAL ? Y : Z;
    const int T = plano == MSL_AXIAL ? Z : Y;
    dim3 grid(((X + kTile - 1) / kTile) * ((W + kTile - 1) / kTile), T, nvol);
    ProfScope prof(K_RECON_GATHER, stream);
    recon_gather_kernel<<<grid, 256, 0, stream>>>(a);
    MSL_LAUNCH_CHECK("recon_gather_kernel");
    return MSL_OK;
}

int launch_consensus_eval(const uint8_t* ax, const uint8_t* co, const uint8_t* sa, const uint8_t* gt,
                          int nvol, size_t nvox, int umbral, uint8_t* consenso, long long* counts, cudaStream_t stream) {
    if (counts) MSL_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)nvol * 16 * sizeof(long long), stream));
    VoteArgs a;
    a.ax = ax; a.co = co; a.sa = sa; a.gt = gt; a.consenso = consenso; a.counts = counts; a.nvox = nvox; a.umbral = umbral;
    dim3 grid(chunks_for(nvox, nvol), nvol);
    const bool vec = (nvox % 8 == 0) && aligned8(ax) && aligned8(co) && aligned8(sa) && aligned8(gt) && aligned8(consenso);
    ProfScope prof(K_CONSENSUS_EVAL, stream);
    if (vec) consensus_eval_kernel<true><<<grid, kCntThreads, 0, stream>>>(a);
    else consensus_eval_kernel<false><<<grid, kCntThreads, 0, stream>>>(a);
    MSL_LAUNCH_CHECK("consensus_eval_kernel");
    return MSL_OK;
}

int launch_confusion_counts(const uint8_t* gt, const uint8_t* pred, int nvol, size_t nvox, long long* counts,
                            cudaStream_t stream) {
    MSL_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)nvol * 4 * sizeof(long long), stream));
    dim3 grid(chunks_for(nvox, nvol), nvol);
    const bool vec = (nvox % 8 == 0) && aligned8(gt) && aligned8(pred);
    ProfScope prof(K_CONFUSION_COUNTS, stream);
    if (vec) confusion_counts_kernel<true><<<grid, kCntThreads, 0, stream>>>(gt, pred, nvox, counts);
    else confusion_counts_kernel<false><<<grid, kCntThreads, 0, stream>>>(gt, pred, nvox, counts);
    MSL_LAUNCH_CHECK("confusion_counts_kernel");
    return MSL_OK;
}

}  // namespace msl
