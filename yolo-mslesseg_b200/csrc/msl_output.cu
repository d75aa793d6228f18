// Output side of the voxel path: prediction post-processing (R0), reconstruction (R1-R2), tri-planar vote (R3),
// confusion counts (R4; per volume and per slice).
//
//  combine_predictions  YOLO instance masks -> one {0, 255} mask per slice (scripts/generar_predicciones.py:123-140).
//  recon          predicted 2-D masks (slice orientation, uint8, pixel > 0 == lesion) -> [Z][Y][X] volumes.
//                 Reference: scripts/reconstruir_volumen.py:146-148 (binarise), :179-186 (insertar_corte),
//                 :199-213 (zero volume + loop).  An inverse slice map resolves duplicate indices.  Axial / coronal:
//                 memset + one CTA per present slice, which assembles the transposed slice in zero-filled shared memory
//                 from the non-zero bytes only and streams it out with 128-bit stores.  Sagital (the slice index is the
//                 volume's fastest axis): the kernel writes the whole volume densely, two slice rows per CTA.
//                 Byte-granular kernels remain as the fallback (float volumes, odd sizes, unaligned buffers).
//  consensus_eval (ax + co + sa >= umbral) fused with the 4x4 confusion counts of the three planes and the
//                 consensus against the ground truth.  Reference: scripts/generar_consenso.py:106-109 and
//                 the boolean sums of utils/utils.py:455-495.  SIMD-within-a-register byte predicates,
//                 popc, warp shuffles, one int64 atomic per counter per CTA.
//  slice_counts   the same counts for every slice of the three planes (extras/visualizar_prediccion_corte.py:150-182).
#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

namespace {

// ------------------------------------------------------------------------------------ combine predictions
// P[r][c] = OR_i (mask_i[sy(r)][sx(c)] > 0.5), (height, width) = (cols, rows); cv2.resize INTER_NEAREST picks
// sx = min(floor(c * ifx), mw - 1) with ifx = 1 / (width / mw) evaluated in double (OpenCV resizeNN), same for y.
// grid (tiles_c, tiles_r, nslices); block 256 = 32 (c) x 8; a CTA covers 32 x 32 pixels of P.  For the slice-oriented
// layout G[a][b] = 255 * P[cols-1-b][a] the tile is transposed through shared memory so both sides stay coalesced.
__global__ void __launch_bounds__(256) combine_predictions_kernel(const float* __restrict__ masks, const int32_t* __restrict__ inst_offset,
                                                                  int mh, int mw, int rows, int cols, int layout,
                                                                  uint8_t* __restrict__ out) {
    __shared__ uint8_t tile[32][33];
    __shared__ int s_sx[32], s_sy[32];
    const int H = cols, Wd = rows;                        // P is H x Wd
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, s = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x < 32) {
        const double ifx = 1.0 / ((double)Wd / (double)mw);
        s_sx[threadIdx.x] = min((int)floor((double)(c0 + threadIdx.x) * ifx), mw - 1);
    } else if (threadIdx.x < 64) {
        const double ify = 1.0 / ((double)H / (double)mh);
        s_sy[threadIdx.x - 32] = min((int)floor((double)(r0 + threadIdx.x - 32) * ify), mh - 1);
    }
    __syncthreads();
    const int i0 = inst_offset[s], i1 = inst_offset[s + 1];
    const size_t msz = (size_t)mh * mw;
    const int c = c0 + tx;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int rl = ty + 8 * k, r = r0 + rl;
        uint8_t v = 0;
        if (r < H && c < Wd) {
            const float* m = masks + (size_t)i0 * msz + (size_t)s_sy[rl] * mw + s_sx[tx];
            for (int i = i0; i < i1; ++i, m += msz)
                if (__ldg(m) > 0.5f) { v = 1; break; }
        }
        if (layout == MSL_OUT_P) { if (r < H && c < Wd) out[((size_t)s * H + r) * Wd + c] = v; }
        else tile[rl][tx] = v;
    }
    if (layout == MSL_OUT_P) return;
    __syncthreads();
    // G[a][b], a = c (rows), b = cols - 1 - r: lanes run along b (descending r)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int cl = ty + 8 * k, a = c0 + cl;
        const int r = r0 + 31 - tx;                       // lane 0 takes the tile's last P row = the smallest b
        if (a < Wd && r < H) out[((size_t)s * rows + a) * cols + (cols - 1 - r)] = tile[31 - tx][cl] ? 255 : 0;
    }
}

// ------------------------------------------------------------------------------------ recon
// Volumes are zero-filled by cudaMemsetAsync; only the slices that exist are written (typically ~20 % of the
// indices of a plane: Paciente.indices_a_usar keeps a central window of the lesion slices).
__global__ void recon_init_kernel(int32_t* slot_of, size_t nmap, int32_t* xrange, int nvol) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nmap) slot_of[i] = -1;
    if (i < (size_t)nvol) { xrange[2 * i] = 0x7fffffff; xrange[2 * i + 1] = -1; }
}

__global__ void slot_map_kernel(const int32_t* __restrict__ vol_of_slice, const int32_t* __restrict__ idx_of_slice,
                                int nslices, int nvol, int n_plane, int32_t* __restrict__ slot_of, int32_t* __restrict__ xrange) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslices) return;
    int v = vol_of_slice[s], i = idx_of_slice[s];
    if (v < 0 || v >= nvol || i < 0 || i >= n_plane) return;
    atomicMax(&slot_of[(size_t)v * n_plane + i], s);     // duplicate index: the later slice wins
    atomicMin(&xrange[2 * v], i);
    atomicMax(&xrange[2 * v + 1], i);
}

struct ReconArgs {
    const uint8_t* slices;
    size_t slice_pitch;
    const int32_t* vol_of_slice;
    const int32_t* idx_of_slice;
    const int32_t* slot_of;
    const int32_t* xrange;
    uint8_t* vol_u8;
    float* vol_f32;
    int X, Y, Z, plano, nvol;
};

// Axial / coronal: one CTA per PRESENT slice.  The slice Q[x][w] (w = y for axial, z for coronal; w fastest) is
// binarised and transposed into shared memory, then written as x-contiguous rows of the volume:
//   axial   slice k: out[v][k][w][x]     coronal slice j: out[v][w][j][x]
// grid (nslices); block 256.  PAIR: 16-bit accesses (X and the slice row length even, 2-byte aligned bases).
template <bool PAIR>
__global__ void __launch_bounds__(256) recon_rows_kernel(const ReconArgs a) {
    extern __shared__ __align__(16) uint8_t T[];          // [W][tp]  transposed, binarised slice
    const int X = a.X, Y = a.Y, Z = a.Z;
    const int s = blockIdx.x;
    const int v = a.vol_of_slice[s], idx = a.idx_of_slice[s];
    const int n_plane = a.plano == MSL_AXIAL ? Z : Y;
    if (v < 0 || v >= a.nvol || idx < 0 || idx >= n_plane) return;
    if (a.slot_of[(size_t)v * n_plane + idx] != s) return;          // superseded duplicate
    const int W = a.plano == MSL_AXIAL ? Y : Z;           // slice row length (cols), number of output rows
    const int tp = (X + 4) & ~1;                          // smem pitch: even and not a multiple of 4 words apart
    const uint8_t* src = a.slices + (size_t)s * a.slice_pitch;
    const int tid = threadIdx.x;
    if (PAIR) {
        const uint16_t* src2 = reinterpret_cast<const uint16_t*>(src);
        const int wp = W >> 1;
        for (int q = tid; q < X * wp; q += 256) {
            const int x = q / wp, w = 2 * (q - x * wp);
            const uint32_t u = __ldg(src2 + q);
            T[w * tp + x] = (u & 0xff) ? 1 : 0;
            T[(w + 1) * tp + x] = (u >> 8) ? 1 : 0;
        }
    } else {
        for (int q = tid; q < X * W; q += 256) {
            const int x = q / W, w = q - x * W;
            T[w * tp + x] = __ldg(src + q) ? 1 : 0;
        }
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int w = warp; w < W; w += 8) {
        const size_t off = a.plano == MSL_AXIAL ? (((size_t)v * Z + idx) * Y + w) * X : (((size_t)v * Z + w) * Y + idx) * X;
        if (a.vol_u8) {
            if (PAIR) {
                uint16_t* dst = reinterpret_cast<uint16_t*>(a.vol_u8 + off);
                const uint16_t* row = reinterpret_cast<const uint16_t*>(T + w * tp);
                for (int j = lane; j < (X >> 1); j += 32) dst[j] = row[j];
            } else {
                for (int x = lane; x < X; x += 32) a.vol_u8[off + x] = T[w * tp + x];
            }
        }
        if (a.vol_f32)
            for (int x = lane; x < X; x += 32) a.vol_f32[off + x] = (float)T[w * tp + x];
    }
}

// Sagital: out[v][z][y][x] = Q_x[y][z] - the slice index is the volume's FASTEST axis, so slices are gathered
// through 64 (x) x 64 (z) byte tiles; only tiles that overlap the range of present slices do any work.
constexpr int kTile = 64;
constexpr int kTilePitch = 68;     // 17 words: conflict-free column reads

// grid (tiles_z, Y, nvol); block 256.  Each CTA walks the x-tiles that overlap [first, last] present slice.
__global__ void __launch_bounds__(256) recon_sagital_kernel(const ReconArgs a) {
    __shared__ __align__(4) uint8_t tile[kTile][kTilePitch];
    const int X = a.X, Y = a.Y, Z = a.Z;
    const int z0 = blockIdx.x * kTile;
    const int y = blockIdx.y, v = blockIdx.z;
    const int xmin = a.xrange[2 * v], xmax = a.xrange[2 * v + 1];
    if (xmax < 0) return;                                 // this volume has no slice at all
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t* slot_v = a.slot_of + (size_t)v * X;
    const bool even = ((a.slice_pitch | (size_t)Z | reinterpret_cast<uintptr_t>(a.slices)) & 1) == 0;
    const bool even_out = (X & 1) == 0;
    for (int x0 = (xmin / kTile) * kTile; x0 <= xmax; x0 += kTile) {
        for (int xr = warp; xr < kTile; xr += 8) {
            const int x = x0 + xr;
            const int slot = x < X ? slot_v[x] : -1;
            const uint8_t* src = a.slices + (slot < 0 ? 0 : (size_t)slot * a.slice_pitch) + (size_t)y * Z + z0;
            const int z = z0 + 2 * lane;
            uint32_t q0 = 0, q1 = 0;
            if (slot >= 0) {
                if (even && z + 1 < Z) { const uint32_t u = __ldg(reinterpret_cast<const uint16_t*>(src) + lane); q0 = u & 0xff; q1 = u >> 8; }
                else { if (z < Z) q0 = __ldg(src + 2 * lane); if (z + 1 < Z) q1 = __ldg(src + 2 * lane + 1); }
            }
            *reinterpret_cast<uint16_t*>(&tile[xr][2 * lane]) = (uint16_t)((q0 ? 1u : 0u) | (q1 ? 0x100u : 0u));
        }
        __syncthreads();
        for (int zr = warp; zr < kTile; zr += 8) {
            const int z = z0 + zr;
            if (z >= Z) break;
            const size_t off = (((size_t)v * Z + z) * Y + y) * X + x0;
            const int xr = 2 * lane, x = x0 + xr;
            const uint32_t b0 = tile[xr][zr], b1 = tile[xr + 1][zr];
            if (a.vol_u8) {
                if (even_out && x + 1 < X) *reinterpret_cast<uint16_t*>(a.vol_u8 + off + xr) = (uint16_t)(b0 | (b1 << 8));
                else { if (x < X) a.vol_u8[off + xr] = (uint8_t)b0; if (x + 1 < X) a.vol_u8[off + xr + 1] = (uint8_t)b1; }
            }
            if (a.vol_f32) {
                if (x < X) a.vol_f32[off + xr] = (float)b0;
                if (x + 1 < X) a.vol_f32[off + xr + 1] = (float)b1;
            }
        }
        __syncthreads();
    }
}

// ---- sparse-scatter versions (uint8 volumes, even X, 4-byte aligned slices): the kernels above remain the fallback.
// Lesion masks are almost empty, and a transposed copy of a slice costs ~10 instructions per byte however it is done.
// So the transposed slice is ASSEMBLED in zero-filled shared memory - the slice is scanned with 128-bit loads and only
// its non-zero bytes are scattered - and then streamed to the volume with 128-bit stores.

// Axial / coronal, one CTA per present slice Q[x][w] (w = y for axial, z for coronal; w fastest):
//   axial   slice k: out[v][k][w][0 .. X)     coronal slice j: out[v][w][j][0 .. X)
// Output row w lives at T2 + w * pitch with pitch == (global row stride) (mod 16) and T2 shifted by the alignment of row 0,
// so every row has the same 16-byte phase in shared memory and in the volume: [16-bit head] + 128-bit body + [16-bit tail].
__global__ void __launch_bounds__(256) recon_rows_sparse_kernel(const ReconArgs a, int pitch) {
    extern __shared__ __align__(16) uint8_t Traw[];
    const int X = a.X, Y = a.Y, Z = a.Z;
    const int s = blockIdx.x;
    const int v = a.vol_of_slice[s], idx = a.idx_of_slice[s];
    const int n_plane = a.plano == MSL_AXIAL ? Z : Y;
    if (v < 0 || v >= a.nvol || idx < 0 || idx >= n_plane) return;
    if (a.slot_of[(size_t)v * n_plane + idx] != s) return;          // superseded duplicate
    const int W = a.plano == MSL_AXIAL ? Y : Z;           // slice row length, number of output rows
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t row0 = a.plano == MSL_AXIAL ? (((size_t)v * Z + idx) * Y) * X : (((size_t)v * Z) * Y + idx) * X;
    const size_t rstride = a.plano == MSL_AXIAL ? (size_t)X : (size_t)Y * X;
    uint8_t* g0 = a.vol_u8 + row0;
    uint8_t* T2 = Traw + (reinterpret_cast<uintptr_t>(g0) & 15);
    // loads first (four 128-bit vectors per thread in flight), zero-fill while they are on their way
    const uint8_t* src = a.slices + (size_t)s * a.slice_pitch;
    const int nbytes = X * W;                             // multiple of 4 (launcher)
    const int headb = min(nbytes, (int)((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15));     // multiple of 4
    const int nvec = (nbytes - headb) >> 4, tailb = headb + (nvec << 4);
    const uint4* src4 = reinterpret_cast<const uint4*>(src + headb);
    const unsigned magic_w = (unsigned)(0x100000000ull / (unsigned)W) + 1u;       // W >= 2 (even)
    auto scatter = [&](uint32_t u, int o) {               // the four bytes at flat offset o of the slice
        if (u == 0) return;
        int x = (int)__umulhi((unsigned)o, magic_w), w = o - x * W;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if ((u >> (8 * j)) & 0xffu) T2[w * pitch + x] = 1;
            if (++w == W) { w = 0; ++x; }
        }
    };
    uint4 pre[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int q = tid + 256 * i; pre[i] = q < nvec ? __ldg(src4 + q) : make_uint4(0, 0, 0, 0); }
    {
        uint4* t4 = reinterpret_cast<uint4*>(Traw);
        const int n4 = (W * pitch + 16 + 15) >> 4;
        for (int q = tid; q < n4; q += 256) t4[q] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    auto scan16 = [&](const uint4& u, int q) {
        if ((u.x | u.y | u.z | u.w) == 0) return;
        const int o = headb + 16 * q;
        scatter(u.x, o); scatter(u.y, o + 4); scatter(u.z, o + 8); scatter(u.w, o + 12);
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) scan16(pre[i], tid + 256 * i);
    for (int q0 = 4 * 256 + tid; q0 < nvec; q0 += 4 * 256) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int q = q0 + 256 * i; pre[i] = q < nvec ? __ldg(src4 + q) : make_uint4(0, 0, 0, 0); }
#pragma unroll
        for (int i = 0; i < 4; ++i) scan16(pre[i], q0 + 256 * i);
    }
    for (int o = 4 * tid; o < headb; o += 4 * 256) scatter(__ldg(reinterpret_cast<const uint32_t*>(src + o)), o);
    for (int o = tailb + 4 * tid; o < nbytes; o += 4 * 256) scatter(__ldg(reinterpret_cast<const uint32_t*>(src + o)), o);
    __syncthreads();
    // stream out, a warp per row; 8 * rstride == 0 (mod 16), so a warp's rows all split the same way
    uint8_t* g = g0 + warp * rstride;
    const uint8_t* sm = T2 + warp * pitch;
    const int hb = min(X, (int)((16 - (reinterpret_cast<uintptr_t>(g) & 15)) & 15));              // even
    const int nv = (X - hb) >> 4, tb = hb + 16 * nv;
    const int nhead = hb >> 1, ntail = (X - tb) >> 1;     // halfwords, <= 7 each
    const int e = 31 - lane;                              // lanes 31, 30, ... : head halfwords, then tail halfwords
    const int hoff = e < nhead ? 2 * e : (e - nhead < ntail ? tb + 2 * (e - nhead) : -1);
    for (int w = warp; w < W; w += 8, g += 8 * rstride, sm += 8 * pitch) {
#pragma unroll 1
        for (int q = lane; q < nv; q += 32)
            *reinterpret_cast<uint4*>(g + hb + 16 * q) = *reinterpret_cast<const uint4*>(sm + hb + 16 * q);
        if (hoff >= 0) *reinterpret_cast<uint16_t*>(g + hoff) = *reinterpret_cast<const uint16_t*>(sm + hoff);
    }
}

// Sagital: out[v][z][y][x] = Q_x[y][z] - the slice index is the volume's FASTEST axis, so a present slice contributes one
// byte to every row of the volume.  Scattering those bytes (or 40-byte fragments of them) costs a read-modify-write per
// 32-byte sector, so this kernel writes the volume DENSELY instead (and the launcher skips the memset): one CTA per
// (2 slice rows y0, y0+1; volume) assembles, for every z, the 2 * X contiguous output bytes out[z][y0 .. y0+1][0 .. X) in
// shared memory - zeros, then the bytes of the present slices, transposed on the way in (only ~20 % of the indices are
// present) - and streams them out as aligned 32-bit words.  grid (ceil(Y / 2), nvol); block 512;
// smem Z * pitch + 32 + (2 X + 1) * 4 bytes.
constexpr int kSagRows = 2;
constexpr int kSagThreads = 512;
constexpr int kSagWarps = kSagThreads / 32;
constexpr int kSagBatch = 8;                              // 32-bit loads in flight per lane
__global__ void __launch_bounds__(kSagThreads) recon_sagital_dense_kernel(const ReconArgs a, int pitch) {
    extern __shared__ __align__(16) uint8_t Sraw[];
    const int X = a.X, Y = a.Y, Z = a.Z;
    const int y0 = blockIdx.x * kSagRows, v = blockIdx.y;
    // [Z][pitch]: byte c = yy * X + x of the chunk of plane z; same alignment (mod 16) as the chunk's place in the volume
    uint8_t* S = Sraw + (reinterpret_cast<uintptr_t>(a.vol_u8 + (((size_t)v * Z) * Y + y0) * X) & 15);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ny = min(kSagRows, Y - y0);
    const int chunk = ny * X;                             // bytes per z, a multiple of 4 (launcher)
    int32_t* s_px = reinterpret_cast<int32_t*>(Sraw + (((size_t)Z * pitch + 16 + 15) & ~(size_t)15));   // [X] present slice indices
    int32_t* s_ps = s_px + X;                             // [X] their slots, [X] the count
    {
        uint4* s4 = reinterpret_cast<uint4*>(Sraw);
        const int n4 = (Z * pitch + 16 + 15) >> 4;        // the launcher's allocation is rounded up to 16 bytes
        for (int q = tid; q < n4; q += kSagThreads) s4[q] = make_uint4(0, 0, 0, 0);
    }
    if (warp == 0) {                                      // compact list of the present slices
        int cnt = 0;
        for (int xb = 0; xb < X; xb += 32) {
            const int x = xb + lane;
            const int slot = x < X ? a.slot_of[(size_t)v * X + x] : -1;
            const unsigned m = __ballot_sync(FULL, slot >= 0);
            if (slot >= 0) { const int i = cnt + __popc(m & ((1u << lane) - 1u)); s_px[i] = x; s_ps[i] = slot; }
            cnt += __popc(m);
        }
        if (lane == 0) s_ps[X] = cnt;
    }
    __syncthreads();
    // present slices: ny * Z contiguous bytes each, read as 32-bit words (task = slice x group of 32 words, warp-uniform
    // slice base, kSagBatch loads per lane in flight).  S is zero-filled and lesion masks are sparse, so only non-zero
    // words cost anything beyond the load (the branch is warp-uniform almost everywhere).
    {
        const int nwz = (ny * Z) >> 2;                    // ny * Z is a multiple of 4 (launcher)
        const int ngrp = (nwz + 31) >> 5, ntask = s_ps[X] * ngrp;
        const unsigned magic_g = (unsigned)(0x100000000ull / (unsigned)ngrp) + 1u;
        const uint8_t* base = a.slices + (size_t)y0 * Z;
        for (int k0 = warp; k0 < ntask; k0 += kSagBatch * kSagWarps) {
            uint32_t u[kSagBatch], meta[kSagBatch];
#pragma unroll
            for (int i = 0; i < kSagBatch; ++i) {
                const int k = k0 + kSagWarps * i;
                u[i] = 0; meta[i] = 0;
                if (k < ntask) {
                    const int xi = ngrp == 1 ? k : (int)__umulhi((unsigned)k, magic_g), w = (k - xi * ngrp) * 32 + lane;
                    if (w < nwz) {
                        u[i] = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)s_ps[xi] * a.slice_pitch) + w);
                        meta[i] = (uint32_t)w | ((uint32_t)s_px[xi] << 16);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kSagBatch; ++i)
                if (u[i]) {
                    const int x = (int)(meta[i] >> 16);
                    int b = 4 * (int)(meta[i] & 0xffffu);
#pragma unroll
                    for (int j = 0; j < 4; ++j, ++b)
                        if ((u[i] >> (8 * j)) & 0xffu) { const int yy = b >= Z ? 1 : 0; S[(b - yy * Z) * pitch + yy * X + x] = 1; }
                }
        }
    }
    __syncthreads();
    // stream out: the chunk of plane z starts at the same offset (mod 16) in shared memory and in the volume (the launcher
    // picks pitch == Y * X (mod 16), the kernel shifts S by the alignment of the first chunk), so its body moves as
    // 128-bit vectors; a warp per plane, the lanes behind the vectors take the unaligned head / tail words.
    uint8_t* out = a.vol_u8 + (((size_t)v * Z) * Y + y0) * X;
    const size_t zstride = (size_t)Y * X;
    uint8_t* g = out + warp * zstride;
    const uint8_t* sm = S + warp * pitch;
    // kSagWarps * zstride == 0 (mod 16): the head / body / tail split of a warp's planes never changes
    const int headb = min(chunk, (int)((16 - (reinterpret_cast<uintptr_t>(g) & 15)) & 15));
    const int nvec = (chunk - headb) >> 4, tailb = headb + 16 * nvec;
    const int nhead = headb >> 2, ntail = (chunk - tailb) >> 2;
    const int e = 31 - lane;                              // lanes 31, 30, ... : head words, then tail words
    const int woff = e < nhead ? 4 * e : (e - nhead < ntail ? tailb + 4 * (e - nhead) : -1);
    if (nvec <= 32) {
        const int voff = lane < nvec ? headb + 16 * lane : -1;
        for (int z = warp; z < Z; z += kSagWarps, g += kSagWarps * zstride, sm += kSagWarps * pitch) {
            if (voff >= 0) *reinterpret_cast<uint4*>(g + voff) = *reinterpret_cast<const uint4*>(sm + voff);
            if (woff >= 0) *reinterpret_cast<uint32_t*>(g + woff) = *reinterpret_cast<const uint32_t*>(sm + woff);
        }
    } else {
        for (int z = warp; z < Z; z += kSagWarps, g += kSagWarps * zstride, sm += kSagWarps * pitch) {
#pragma unroll 1
            for (int q = lane; q < nvec; q += 32)
                *reinterpret_cast<uint4*>(g + headb + 16 * q) = *reinterpret_cast<const uint4*>(sm + headb + 16 * q);
            if (woff >= 0) *reinterpret_cast<uint32_t*>(g + woff) = *reinterpret_cast<const uint32_t*>(sm + woff);
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
// exact predicates of the reference (==1 / ==0 per byte); used for the rare words that hold a non-binary byte
__device__ __forceinline__ void count4(Counts4& c, uint32_t g1, uint32_t g0, uint32_t p) {
    uint32_t p1 = one_bytes(p), p0 = zero_bytes(p);
    c.tp += __popc(g1 & p1); c.fp += __popc(g0 & p1);
    c.fn += __popc(g1 & p0); c.tn += __popc(g0 & p0);
}

// Binary words (every byte 0 or 1 - the only kind real masks contain) are counted with byte-lane SIMD adds on the
// main ALU pipe: per plane only sum(g & p) and sum(p) are accumulated (+ one sum(g)); tp = sum(g&p),
// fp = sum(p) - tp, fn = sum(g) - tp, tn = n - tp - fp - fn.  Byte lanes hold at most kFlush * 2 <= 254.
constexpr int kCntThreads = 256;
constexpr int kFlush = 120;

template <int NPLANE>
struct BinAcc {
    uint32_t gp[NPLANE], p[NPLANE], g;
    int s_gp[NPLANE], s_p[NPLANE], s_g, n;
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < NPLANE; ++i) { gp[i] = 0; p[i] = 0; s_gp[i] = 0; s_p[i] = 0; }
        g = 0; s_g = 0; n = 0;
    }
    __device__ __forceinline__ void flush() {
#pragma unroll
        for (int i = 0; i < NPLANE; ++i) {
            s_gp[i] = __dp4a(gp[i], 0x01010101u, (unsigned)s_gp[i]); gp[i] = 0;
            s_p[i] = __dp4a(p[i], 0x01010101u, (unsigned)s_p[i]); p[i] = 0;
        }
        s_g = __dp4a(g, 0x01010101u, (unsigned)s_g); g = 0;
    }
    __device__ __forceinline__ void fold(Counts4 (&c)[NPLANE]) {
        flush();
#pragma unroll
        for (int i = 0; i < NPLANE; ++i) {
            const int tp = s_gp[i], fp = s_p[i] - tp, fn = s_g - tp;
            c[i].tp += tp; c[i].fp += fp; c[i].fn += fn; c[i].tn += n - tp - fp - fn;
        }
    }
};

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

__device__ __forceinline__ void vote_count_word(uint32_t g, uint32_t a, uint32_t c, uint32_t s, uint32_t r,
                                                BinAcc<4>& acc, Counts4 (&slow)[4]) {
    if (((g | a | c | s) & 0xfefefefeu) == 0) {
        acc.g += g;
        acc.gp[0] += g & a; acc.p[0] += a;
        acc.gp[1] += g & c; acc.p[1] += c;
        acc.gp[2] += g & s; acc.p[2] += s;
        acc.gp[3] += g & r; acc.p[3] += r;
        acc.n += 4;
    } else {
        const uint32_t g1 = one_bytes(g), g0 = zero_bytes(g);
        count4(slow[0], g1, g0, a); count4(slow[1], g1, g0, c); count4(slow[2], g1, g0, s); count4(slow[3], g1, g0, r);
    }
}

// grid (chunks, nvol).  VEC: 8-byte granules (nvox % 8 == 0 and 8-byte aligned bases) or bytes.
template <bool VEC>
__global__ void __launch_bounds__(kCntThreads, 4) consensus_eval_kernel(const VoteArgs a) {
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
        BinAcc<4> acc;
        acc.init();
        int it = 0;
        // two granules per iteration: eight independent 64-bit loads in flight per thread
        const size_t stride = (size_t)gridDim.x * kCntThreads;
        for (size_t g = (size_t)blockIdx.x * kCntThreads + threadIdx.x; g < ng; g += 2 * stride) {
            const size_t g2 = g + stride;
            const bool has2 = g2 < ng;
            const size_t gb = has2 ? g2 : g;
            const uint2 wa = __ldg(ax + g), wc = __ldg(co + g), ws = __ldg(sa + g);
            const uint2 xa = __ldg(ax + gb), xc = __ldg(co + gb), xs = __ldg(sa + gb);
            const uint2 wg = cnt ? __ldg(gt + g) : make_uint2(0, 0);
            const uint2 xg = cnt ? __ldg(gt + gb) : make_uint2(0, 0);
            uint2 r, r2;
            // masks are ~99 % zeros: sixteen background voxels (all four volumes zero, threshold above zero) only count
            if (a.umbral > 0 && (wa.x | wa.y | wc.x | wc.y | ws.x | ws.y | wg.x | wg.y | xa.x | xa.y | xc.x | xc.y | xs.x | xs.y | xg.x | xg.y) == 0) {
                if (out) { out[g] = make_uint2(0, 0); if (has2) out[g2] = make_uint2(0, 0); }
                acc.n += has2 ? 16 : 8;
                continue;
            }
            r.x = vote4(wa.x, wc.x, ws.x, a.umbral);
            r.y = vote4(wa.y, wc.y, ws.y, a.umbral);
            r2.x = vote4(xa.x, xc.x, xs.x, a.umbral);
            r2.y = vote4(xa.y, xc.y, xs.y, a.umbral);
            if (out) { out[g] = r; if (has2) out[g2] = r2; }
            if (cnt) {
                vote_count_word(wg.x, wa.x, wc.x, ws.x, r.x, acc, c);
                vote_count_word(wg.y, wa.y, wc.y, ws.y, r.y, acc, c);
                if (has2) {
                    vote_count_word(xg.x, xa.x, xc.x, xs.x, r2.x, acc, c);
                    vote_count_word(xg.y, xa.y, xc.y, xs.y, r2.y, acc, c);
                }
                if (++it == kFlush / 2) { acc.flush(); it = 0; }
            }
        }
        if (cnt) acc.fold(c);
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
        BinAcc<1> acc;
        acc.init();
        int it = 0;
        auto word = [&](uint32_t g, uint32_t p) {
            if (((g | p) & 0xfefefefeu) == 0) { acc.g += g; acc.gp[0] += g & p; acc.p[0] += p; acc.n += 4; }
            else count4(c[0], one_bytes(g), zero_bytes(g), p);
        };
        for (size_t g = (size_t)blockIdx.x * kCntThreads + threadIdx.x; g < ng; g += (size_t)gridDim.x * kCntThreads) {
            const uint2 wg = __ldg(g8 + g), wp = __ldg(p8 + g);
            if ((wg.x | wg.y | wp.x | wp.y) == 0) { acc.n += 8; continue; }       // background: only counted
            word(wg.x, wp.x);
            word(wg.y, wp.y);
            if (++it == kFlush) { acc.flush(); it = 0; }
        }
        acc.fold(c);
    } else {
        for (size_t i = (size_t)blockIdx.x * kCntThreads + threadIdx.x; i < nvox; i += (size_t)gridDim.x * kCntThreads) {
            uint32_t g = gt[off + i] | 0x02020200u, p = pred[off + i] | 0x02020200u;
            count4(c[0], one_bytes(g), zero_bytes(g), p);
        }
    }
    flush_counts<1>(c, counts + (size_t)v * 4);
}

// Per-slice counts of the three planes in one flat scan.  Masks are ~99 % zeros: a 16-byte vector of gt | pred that is
// zero costs two loads and an OR; every other voxel is classified (tp / fp / fn / neither-0-nor-1) and added to
// shared-memory counters of its three slices, which the CTA flushes once.  tn follows from the slice size.
// grid (ceil(nvec / (8 * kCntThreads)), nvol); dynamic smem (Z + Y + X) * 4 ints; counts pre-zeroed.
constexpr int kSliceVecs = 8;
__global__ void __launch_bounds__(kCntThreads) slice_counts_kernel(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pred,
                                                                   unsigned nvox, int X, int Y, int Z, long long* __restrict__ counts) {
    extern __shared__ int s_cnt[];           // [Z + Y + X][4]: tp, fp, fn, other
    const int v = blockIdx.y, nsl = Z + Y + X;
    const uint8_t* pg = gt + (size_t)v * nvox;
    const uint8_t* pp = pred + (size_t)v * nvox;
    for (int i = threadIdx.x; i < nsl * 4; i += kCntThreads) s_cnt[i] = 0;
    __syncthreads();
    const unsigned npl = (unsigned)X * (unsigned)Y;
    bool any = false;
    auto voxel = [&](uint32_t g, uint32_t p, unsigned z, unsigned y, unsigned x) {
        const int k = (g == 1 && p == 1) ? 0 : (g == 0 && p == 1) ? 1 : (g == 1 && p == 0) ? 2 : 3;
        atomicAdd(&s_cnt[4 * z + k], 1);
        atomicAdd(&s_cnt[4 * (Z + y) + k], 1);
        atomicAdd(&s_cnt[4 * (Z + Y + x) + k], 1);
    };
    // bytes [o, o + n) of the volume, n <= 16
    auto scan = [&](const uint32_t (&wg)[4], const uint32_t (&wp)[4], unsigned o, int n) {
        unsigned z = o / npl, r = o - z * npl, y = r / (unsigned)X, x = r - y * (unsigned)X;
        any = true;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const uint32_t g = (wg[i >> 2] >> (8 * (i & 3))) & 0xffu, p = (wp[i >> 2] >> (8 * (i & 3))) & 0xffu;
            if (g | p) voxel(g, p, z, y, x);
            if (++x == (unsigned)X) { x = 0; if (++y == (unsigned)Y) { y = 0; ++z; } }
        }
    };
    // both volumes share the alignment of the vector body only if their bases agree (mod 16); otherwise bytes
    const bool vec = ((reinterpret_cast<uintptr_t>(pg) ^ reinterpret_cast<uintptr_t>(pp)) & 15) == 0;
    const unsigned head = vec ? min(nvox, (unsigned)((16 - (reinterpret_cast<uintptr_t>(pg) & 15)) & 15)) : nvox;
    const unsigned nvec = (nvox - head) / 16;
    if (vec) {
        const uint4* g4 = reinterpret_cast<const uint4*>(pg + head);
        const uint4* p4 = reinterpret_cast<const uint4*>(pp + head);
        const unsigned q0 = blockIdx.x * (unsigned)(kSliceVecs * kCntThreads) + threadIdx.x;
        uint4 a[kSliceVecs], b[kSliceVecs];
#pragma unroll
        for (int j = 0; j < kSliceVecs; ++j) {
            const unsigned q = q0 + j * kCntThreads;
            a[j] = q < nvec ? __ldg(g4 + q) : make_uint4(0, 0, 0, 0);
            b[j] = q < nvec ? __ldg(p4 + q) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < kSliceVecs; ++j)
            if (a[j].x | a[j].y | a[j].z | a[j].w | b[j].x | b[j].y | b[j].z | b[j].w) {
                const uint32_t wg[4] = {a[j].x, a[j].y, a[j].z, a[j].w}, wp[4] = {b[j].x, b[j].y, b[j].z, b[j].w};
                scan(wg, wp, head + (q0 + j * kCntThreads) * 16, 16);
            }
    }
    if (blockIdx.x == 0) {                   // unaligned head and tail (or everything, if the bases disagree mod 16)
        for (unsigned o = threadIdx.x; o < head; o += kCntThreads)
            if (pg[o] | pp[o]) { const uint32_t wg[4] = {pg[o], 0, 0, 0}, wp[4] = {pp[o], 0, 0, 0}; scan(wg, wp, o, 1); }
        for (unsigned o = head + nvec * 16 + threadIdx.x; o < nvox; o += kCntThreads)
            if (pg[o] | pp[o]) { const uint32_t wg[4] = {pg[o], 0, 0, 0}, wp[4] = {pp[o], 0, 0, 0}; scan(wg, wp, o, 1); }
    }
    if (!__syncthreads_or(any)) return;
    long long* dst = counts + (size_t)v * nsl * 4;
    for (int i = threadIdx.x; i < nsl * 4; i += kCntThreads)
        if (s_cnt[i]) atomicAdd(reinterpret_cast<unsigned long long*>(dst) + i, (unsigned long long)s_cnt[i]);
}

// slot 3 holds the voxels that are neither 0 nor 1 in one of the masks: tn = slice size - tp - fp - fn - other
__global__ void slice_counts_finalize_kernel(long long* counts, int nvol, int X, int Y, int Z) {
    const int nsl = Z + Y + X;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nvol * nsl) return;
    const int sidx = (int)(i % nsl);
    const long long n = sidx < Z ? (long long)X * Y : (sidx < Z + Y ? (long long)X * Z : (long long)Y * Z);
    long long* c = counts + i * 4;
    c[3] = n - c[0] - c[1] - c[2] - c[3];
}

// Which slices (a) and which rows (b) of a uint8 stack [nvol][A][B][C] hold a non-zero byte: the bounding box that a
// host copy needs (everything outside it is zero).  grid (A, nvol); a warp per row, 32-bit loads where the row allows.
// any_a [nvol][A] and any_b [nvol][B] are pre-zeroed; every writer stores 1.
__global__ void __launch_bounds__(256) nonzero_flags_kernel(const uint8_t* __restrict__ stack, int A, int B, int C,
                                                            uint8_t* __restrict__ any_a, uint8_t* __restrict__ any_b) {
    const int ia = blockIdx.x, v = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint8_t* sl = stack + ((size_t)v * A + ia) * (size_t)B * C;
    bool any = false;
    for (int b = warp; b < B; b += 8) {
        const uint8_t* row = sl + (size_t)b * C;
        uint32_t acc = 0;
        const int head = min(C, (int)((4 - (reinterpret_cast<uintptr_t>(row) & 3)) & 3));
        const int nw = (C - head) >> 2;
        const uint32_t* row4 = reinterpret_cast<const uint32_t*>(row + head);
        for (int q = lane; q < nw; q += 32) acc |= __ldg(row4 + q);
        for (int o = lane; o < head; o += 32) acc |= row[o];
        for (int o = head + 4 * nw + lane; o < C; o += 32) acc |= row[o];
        if (__any_sync(FULL, acc != 0)) {
            any = true;
            if (lane == 0) any_b[(size_t)v * B + b] = 1;
        }
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) any_a[(size_t)v * A + ia] = 1;
}

inline bool aligned8(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

inline int chunks_for(size_t nvox, int nvol) {
    // ~8 resident-CTA waves over 148 SMs in total: long grid-stride loops (loads stay in flight) instead of many
    // short-lived CTAs; never less than ~2 granules per thread.
    size_t want = (size_t)(8 * 148 + nvol - 1) / (size_t)nvol;
    size_t maxc = (nvox / 8 + (size_t)kCntThreads * 2 - 1) / ((size_t)kCntThreads * 2);
    if (want > maxc) want = maxc;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace

int launch_recon(const uint8_t* slices, size_t slice_pitch, const int32_t* vol_of_slice, const int32_t* idx_of_slice,
                 int nslices, int plano, int nvol, int X, int Y, int Z, uint8_t* vol_u8, float* vol_f32,
                 int32_t* slot_of, cudaStream_t stream) {
    const int n_plane = plano == MSL_AXIAL ? Z : (plano == MSL_CORONAL ? Y : X);
    const size_t nmap = (size_t)nvol * n_plane;
    const size_t N = (size_t)nvol * X * Y * Z;
    int32_t* xrange = slot_of + nmap;                     // [nvol][2] first / last present index
    // word-granular kernels: uint8 volumes, even X, 4-byte aligned slices whose size is a multiple of 4
    const bool words = nslices > 0 && vol_u8 && !vol_f32 && (X & 1) == 0 && (((size_t)X * Y * Z) < 0x7fffffffull) &&
                       ((slice_pitch | reinterpret_cast<uintptr_t>(slices)) & 3) == 0;
    // dense sagital kernel: every byte of the volume is written by the kernel itself, no memset
    int sag_pitch = 0;
    size_t sag_smem = 0;
    bool sag_dense = words && plano == MSL_SAGITAL && (Z & 1) == 0 && X < 65536 && (size_t)kSagRows * Z < 4 * 65536 && ((size_t)X * Y) % 4 == 0 && (reinterpret_cast<uintptr_t>(vol_u8) & 3) == 0 &&
                     (Y % kSagRows == 0 || X % 4 == 0);
    if (sag_dense) {
        sag_pitch = kSagRows * X;                         // multiple of 4, and == Y * X (mod 16): see the kernel's stream-out
        sag_pitch += (int)(((size_t)X * Y - (size_t)sag_pitch) & 15);
        sag_smem = (((size_t)Z * sag_pitch + 16 + 15) & ~(size_t)15) + (size_t)(2 * X + 1) * sizeof(int32_t);
        if (sag_smem > 227 * 1024) sag_dense = false;
    }
    {
        ProfScope prof(K_RECON_FILL, stream);             // zero fill (not needed by the dense sagital kernel) + map reset
        if (vol_u8 && !sag_dense) MSL_CUDA_CHECK(cudaMemsetAsync(vol_u8, 0, N, stream));
        if (vol_f32) MSL_CUDA_CHECK(cudaMemsetAsync(vol_f32, 0, N * sizeof(float), stream));
        if (nslices <= 0) return MSL_OK;
        recon_init_kernel<<<(unsigned)((nmap + 255) / 256), 256, 0, stream>>>(slot_of, nmap, xrange, nvol);
    }
    MSL_LAUNCH_CHECK("recon_init_kernel");
    {
        ProfScope prof(K_RECON_SLOT_MAP, stream);
        slot_map_kernel<<<(nslices + 255) / 256, 256, 0, stream>>>(vol_of_slice, idx_of_slice, nslices, nvol, n_plane, slot_of, xrange);
    }
    MSL_LAUNCH_CHECK("slot_map_kernel");
    ReconArgs a;
    a.slices = slices; a.slice_pitch = slice_pitch; a.vol_of_slice = vol_of_slice; a.idx_of_slice = idx_of_slice;
    a.slot_of = slot_of; a.xrange = xrange; a.vol_u8 = vol_u8; a.vol_f32 = vol_f32;
    a.X = X; a.Y = Y; a.Z = Z; a.plano = plano; a.nvol = nvol;
    ProfScope prof(K_RECON_GATHER, stream);
    if (sag_dense) {
        MSL_CUDA_CHECK(cudaFuncSetAttribute(recon_sagital_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sag_smem));
        dim3 grid((Y + kSagRows - 1) / kSagRows, nvol);
        recon_sagital_dense_kernel<<<grid, kSagThreads, sag_smem, stream>>>(a, sag_pitch);
        MSL_LAUNCH_CHECK("recon_sagital_dense_kernel");
        return MSL_OK;
    }
    if (words && plano != MSL_SAGITAL && (reinterpret_cast<uintptr_t>(vol_u8) & 1) == 0) {
        const int W = plano == MSL_AXIAL ? Y : Z;
        const size_t rstride = plano == MSL_AXIAL ? (size_t)X : (size_t)X * Y;
        const int pitch = X + (int)((rstride - (size_t)X) & 15);       // == rstride (mod 16), even
        const size_t smem = (((size_t)W * pitch + 16 + 15) & ~(size_t)15);
        if ((W & 1) == 0 && ((size_t)X * W) % 4 == 0 && smem <= 227 * 1024) {
            MSL_CUDA_CHECK(cudaFuncSetAttribute(recon_rows_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            recon_rows_sparse_kernel<<<nslices, 256, smem, stream>>>(a, pitch);
            MSL_LAUNCH_CHECK("recon_rows_sparse_kernel");
            return MSL_OK;
        }
    }
    if (plano == MSL_SAGITAL) {
        dim3 grid((Z + kTile - 1) / kTile, Y, nvol);
        recon_sagital_kernel<<<grid, 256, 0, stream>>>(a);
    } else {
        const int W = plano == MSL_AXIAL ? Y : Z;
        const size_t smem = (size_t)W * ((X + 4) & ~1);
        if (smem > 227 * 1024) { set_error("recon: a %d x %d slice does not fit shared memory", X, W); return MSL_ERR_UNSUPPORTED; }
        const bool pair = ((X | W) & 1) == 0 && ((slice_pitch | reinterpret_cast<uintptr_t>(slices) | reinterpret_cast<uintptr_t>(vol_u8)) & 1) == 0;
        if (pair) {
            MSL_CUDA_CHECK(cudaFuncSetAttribute(recon_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            recon_rows_kernel<true><<<nslices, 256, smem, stream>>>(a);
        } else {
            MSL_CUDA_CHECK(cudaFuncSetAttribute(recon_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            recon_rows_kernel<false><<<nslices, 256, smem, stream>>>(a);
        }
    }
    MSL_LAUNCH_CHECK("recon kernel");
    return MSL_OK;
}

int launch_combine_predictions(const float* masks, const int32_t* inst_offset, int nslices, int mh, int mw, int rows, int cols,
                               int layout, uint8_t* out, cudaStream_t stream) {
    dim3 grid((rows + 31) / 32, (cols + 31) / 32, nslices);
    ProfScope prof(K_COMBINE_PRED, stream);
    combine_predictions_kernel<<<grid, 256, 0, stream>>>(masks, inst_offset, mh, mw, rows, cols, layout, out);
    MSL_LAUNCH_CHECK("combine_predictions_kernel");
    return MSL_OK;
}

int launch_slice_counts(const uint8_t* gt, const uint8_t* pred, int nvol, int X, int Y, int Z, long long* counts, cudaStream_t stream) {
    const int nsl = X + Y + Z;
    const unsigned long long nvox = (unsigned long long)X * Y * Z;
    MSL_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)nvol * nsl * 4 * sizeof(long long), stream));
    const size_t smem = (size_t)nsl * 4 * sizeof(int);
    ProfScope prof(K_SLICE_COUNTS, stream);
    MSL_CUDA_CHECK(cudaFuncSetAttribute(slice_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned per_cta = 16u * kSliceVecs * kCntThreads;
    dim3 grid((unsigned)((nvox + per_cta - 1) / per_cta), nvol);
    slice_counts_kernel<<<grid, kCntThreads, smem, stream>>>(gt, pred, (unsigned)nvox, X, Y, Z, counts);
    MSL_LAUNCH_CHECK("slice_counts_kernel");
    const size_t n = (size_t)nvol * nsl;
    slice_counts_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(counts, nvol, X, Y, Z);
    MSL_LAUNCH_CHECK("slice_counts_finalize_kernel");
    return MSL_OK;
}

int launch_nonzero_flags(const uint8_t* stack, int nvol, int A, int B, int C, uint8_t* any_a, uint8_t* any_b, cudaStream_t stream) {
    MSL_CUDA_CHECK(cudaMemsetAsync(any_a, 0, (size_t)nvol * A, stream));
    MSL_CUDA_CHECK(cudaMemsetAsync(any_b, 0, (size_t)nvol * B, stream));
    ProfScope prof(K_NONZERO_FLAGS, stream);
    nonzero_flags_kernel<<<dim3(A, nvol), 256, 0, stream>>>(stack, A, B, C, any_a, any_b);
    MSL_LAUNCH_CHECK("nonzero_flags_kernel");
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
