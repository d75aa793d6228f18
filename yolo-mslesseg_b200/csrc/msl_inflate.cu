// Byte-stream codec on the device, decode side: RFC 1951 inflate (stored, fixed and dynamic blocks)
// of many independent streams - the .nii.gz volumes behind nib.load (utils/Paciente.py:168, utils/utils.py:156) and the
// zlib streams inside the predicted-mask PNGs behind Image.open / cv2.imread (scripts/reconstruir_volumen.py:141,
// utils/utils.py:391) - so that the host-to-device copies carry file bytes.
//
// inflate_kernel: ONE WARP per stream.  Huffman decoding is serial by nature, so parallelism comes from the number of
//   streams (a .nii.gz written by msl_deflate_chunks is ~450 independent members per volume; a patient has hundreds of
//   mask PNGs).  Every lane runs the same decode loop on the same bit buffer (loads and table look-ups are broadcasts, no
//   divergence); literals are collected one per lane and stored 32 at a time; a match is copied by the whole warp
//   (out[pos + i] = out[pos - dist + i mod dist]).  Decoding goes through a 10-bit (literal / length) and an 8-bit
//   (distance) look-up table in shared memory, longer codes through the canonical count / symbol walk.
// png_unfilter_kernel: PNG scanline filters 0-4 (None, Sub, Up, Average, Paeth) undone by one warp per image, then the
//   first channel is written out as the uint8 mask image msl_recon consumes.
// nifti_convert_kernel: voxel payload of a NIfTI-1 file (uint8 / int16 / int32 / float32 / float64, little endian) -> the
//   float32 or uint8 volume the kernels consume, with the count of values the conversion could not represent exactly.
#include <cstring>

#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

namespace {

constexpr int kInfWarps = 4;
constexpr int kLitBits = 10, kDistBits = 8;
constexpr int kInRing = 64;             // words (two blocks of 32: one being consumed, one prefetched)
constexpr int kOutRing = 2048;          // bytes; matches that reach at most kRingKeep bytes back are served from shared memory
constexpr int kFlush = 512;             // output leaves the ring in blocks of kFlush bytes
constexpr int kRingKeep = kOutRing - 258 - 34;

struct WarpTables {
    uint16_t lit[1 << kLitBits];      // (symbol << 4) | code length; 0 = code longer than kLitBits
    uint16_t dist[1 << kDistBits];    // also the code-length code's table while a dynamic header is read
    uint16_t lit_sorted[288], dist_sorted[32];
    uint16_t codes[320];
    int lit_count[16], dist_count[16];
    int next_code[16], offs[16];
    uint8_t lens[384];                // [0, 19): code-length code; [32, 32 + HLIT + HDIST): literal / length and distance lengths
    uint32_t in_ring[kInRing];        // staged input words: word w of the source lives at in_ring[w % kInRing]
    uint8_t out_ring[kOutRing];       // the last kOutRing output bytes: byte p lives at out_ring[p % kOutRing]
};

struct InfArgs {
    const uint8_t* src;
    const unsigned long long* src_off;   // [n + 1]
    uint8_t* dst;
    const unsigned long long* dst_off;   // [n + 1]
    uint32_t* status;                    // [n][4]: error, bytes produced, stored checksum (Adler-32 / CRC-32 of the last member), stored ISIZE
    int n, container;
    unsigned long long src_bytes;        // readable bytes at src (loads are clamped to it)
    // gzip members this library writes for an all-zero 16 KB chunk (one per match distance): a member whose bytes equal one of
    // them decodes to zeros without being decoded
    uint32_t tmpl[4][kZeroTmplBytes / 4];
    uint32_t tmpl_size[4], tmpl_chk[4], tmpl_raw;
};

enum { INF_OK = 0, INF_ERR_HEADER = 1, INF_ERR_BLOCK = 2, INF_ERR_CODE = 3, INF_ERR_DIST = 4, INF_ERR_SPACE = 5, INF_ERR_INPUT = 6 };

// Bit reader over the warp's staged input.  The compressed bytes are read sequentially, so the warp copies them from
// global to shared memory 32 words at a time (one coalesced load, its latency paid once per 128 bytes instead of once per
// word on the serial decode path); refill() then costs a shared-memory read.
struct BitReader {
    const uint32_t* w;              // src as aligned words
    uint32_t* ring;
    unsigned long long nwords;      // words that may be read
    unsigned long long next;        // index of the next word to move into the bit buffer
    unsigned long long staged;      // words [.., staged) are in the ring
    unsigned long long bb;
    int nb, lane;
    __device__ __forceinline__ void stage() {            // all lanes; keeps at least 8 words ahead of `next`
        if (next + 8 >= staged) {
            const unsigned long long i = staged + lane;
            ring[i % kInRing] = i < nwords ? __ldg(w + i) : 0u;
            staged += 32;
            __syncwarp();
        }
    }
    __device__ __forceinline__ void init(const uint8_t* src, unsigned long long src_bytes, unsigned long long bytepos, uint32_t* ring_, int lane_) {
        w = reinterpret_cast<const uint32_t*>(src);
        ring = ring_; lane = lane_;
        nwords = (src_bytes + 3) >> 2;
        next = bytepos >> 2;
        staged = next;
        __syncwarp();
        stage();
        bb = 0; nb = 0;
        refill();
        const int drop = (int)(bytepos & 3) * 8;
        bb >>= drop; nb -= drop;
        refill();
    }
    __device__ __forceinline__ void refill() {
        if (nb <= 32) {
            stage();
            bb |= (unsigned long long)ring[next % kInRing] << nb;
            nb += 32; ++next;
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(int n) { bb >>= n; nb -= n; }
    __device__ __forceinline__ uint32_t take(int n) { const uint32_t v = peek(n); drop(n); return v; }
    __device__ __forceinline__ unsigned long long bitpos() const { return next * 32ull - (unsigned long long)nb; }
};

// canonical Huffman tables from code lengths (all lanes call it; lane 0 does the serial part)
__device__ void build_table(const uint8_t* lens, int n, uint16_t* fast, int fastbits, int* count, uint16_t* sorted, uint16_t* codes,
                            int* next_code, int* offs, int lane) {
    for (int i = lane; i < (1 << fastbits); i += 32) fast[i] = 0;
    if (lane < 16) count[lane] = 0;
    __syncwarp();
    for (int s = lane; s < n; s += 32) if (lens[s]) atomicAdd(&count[lens[s]], 1);
    __syncwarp();
    if (lane == 0) {
        int code = 0, o = 0;
        for (int b = 1; b <= 15; ++b) { code = (code + (b > 1 ? count[b - 1] : 0)) << 1; next_code[b] = code; offs[b] = o; o += count[b]; }
        for (int s = 0; s < n; ++s) {
            const int l = lens[s];
            if (l) { codes[s] = (uint16_t)next_code[l]++; sorted[offs[l]++] = (uint16_t)s; }
        }
    }
    __syncwarp();
    for (int s = lane; s < n; s += 32) {
        const int l = lens[s];
        if (l && l <= fastbits) {
            const uint32_t rev = __brev((uint32_t)codes[s]) >> (32 - l);
            const uint16_t e = (uint16_t)((s << 4) | l);
            for (uint32_t k = rev; k < (1u << fastbits); k += 1u << l) fast[k] = e;
        }
    }
    __syncwarp();
}

// the fixed code of RFC 1951 section 3.2.6 needs no canonical construction: every look-up entry follows from its index
__device__ void build_fixed_tables(uint16_t* lit, uint16_t* dist, int* lit_count, int* dist_count, int lane) {
    for (int k = lane; k < (1 << kLitBits); k += 32) {
        const uint32_t c7 = __brev((uint32_t)k & 0x7fu) >> 25, c8 = __brev((uint32_t)k & 0xffu) >> 24, c9 = __brev((uint32_t)k & 0x1ffu) >> 23;
        uint32_t sym, l;
        if (c7 <= 0x17u) { sym = 256u + c7; l = 7; }
        else if (c8 >= 0x30u && c8 <= 0xbfu) { sym = c8 - 0x30u; l = 8; }
        else if (c8 >= 0xc0u && c8 <= 0xc7u) { sym = 280u + (c8 - 0xc0u); l = 8; }
        else { sym = 144u + (c9 - 0x190u); l = 9; }
        lit[k] = (uint16_t)((sym << 4) | l);
    }
    for (int k = lane; k < (1 << kDistBits); k += 32) dist[k] = (uint16_t)(((__brev((uint32_t)k & 31u) >> 27) << 4) | 5u);
    if (lane < 16) { lit_count[lane] = 0; dist_count[lane] = 0; }
    __syncwarp();
}

// decode one symbol; returns -1 when no code matches
__device__ __forceinline__ int decode_sym(BitReader& br, const uint16_t* fast, int fastbits, const int* count, const uint16_t* sorted) {
    const uint32_t e = fast[br.peek(fastbits)];
    if (e) { br.drop((int)(e & 15u)); return (int)(e >> 4); }
    // canonical walk (codes longer than the look-up table)
    int code = 0, first = 0, index = 0;
    unsigned long long b = br.bb;
    for (int len = 1; len <= 15; ++len) {
        code |= (int)(b & 1ull); b >>= 1;
        const int c = count[len];
        if (code - c < first) { br.drop(len); return sorted[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

__constant__ uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__global__ void __launch_bounds__(kInfWarps * 32) inflate_kernel(const InfArgs a) {
    __shared__ WarpTables tabs[kInfWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * kInfWarps + warp;
    if (s >= a.n) return;
    WarpTables& T = tabs[warp];
    const unsigned long long sbeg = a.src_off[s], send = a.src_off[s + 1];
    uint8_t* out = a.dst + a.dst_off[s];
    const unsigned long long cap = a.dst_off[s + 1] - a.dst_off[s];
    unsigned long long pos = 0, flushed = 0;          // bytes produced / bytes that left the ring for global memory
    int err = INF_OK;
    uint32_t stored_chk = 0, stored_isize = 0;
    unsigned long long bytepos = sbeg;
    uint8_t* ring = T.out_ring;
    const bool out_aligned = (reinterpret_cast<uintptr_t>(out) & 3) == 0;

    // Output goes to the ring first and leaves it in blocks of kFlush bytes (coalesced 32-bit stores).  The ring keeps the
    // last kOutRing bytes, flushed or not, so near matches never touch global memory.
    auto flush_blocks = [&]() {
        while (pos - flushed >= kFlush) {
            __syncwarp();
            const unsigned mis = (unsigned)(flushed & 3);
            if (mis || !out_aligned) {
                // realign (after a stored block the flushed count is arbitrary); unaligned outputs go byte by byte
                const unsigned nbyte = out_aligned ? 4u - mis : (unsigned)kFlush;
                for (unsigned k = lane; k < nbyte; k += 32) out[flushed + k] = ring[(flushed + k) % kOutRing];
                flushed += nbyte;
                continue;
            }
            const uint32_t* r32 = reinterpret_cast<const uint32_t*>(ring);
            uint32_t* o32 = reinterpret_cast<uint32_t*>(out + flushed);
            const unsigned w0 = (unsigned)(flushed % kOutRing) >> 2;
#pragma unroll
            for (int k = 0; k < kFlush / 128; ++k) o32[k * 32 + lane] = r32[(w0 + k * 32 + lane) % (kOutRing / 4)];
            flushed += kFlush;
        }
    };
    auto flush_all = [&]() {
        flush_blocks();
        __syncwarp();
        if (out_aligned && (flushed & 3) == 0) {
            const unsigned nw = (unsigned)(pos - flushed) >> 2;
            const uint32_t* r32 = reinterpret_cast<const uint32_t*>(ring);
            uint32_t* o32 = reinterpret_cast<uint32_t*>(out + flushed);
            const unsigned w0 = (unsigned)(flushed % kOutRing) >> 2;
            for (unsigned w = lane; w < nw; w += 32) o32[w] = r32[(w0 + w) % (kOutRing / 4)];
            flushed += 4ull * nw;
        }
        for (unsigned long long k = flushed + lane; k < pos; k += 32) out[k] = ring[k % kOutRing];
        flushed = pos;
        __syncwarp();
    };

    if (a.container == MSL_Z_GZIP && a.tmpl_raw && cap >= a.tmpl_raw) {
        const unsigned long long msize = send - sbeg;
        for (int k = 0; k < 4; ++k) {
            if (msize != a.tmpl_size[k]) continue;                                  // (warp-uniform)
            bool eq = true;
            const uint8_t* t = reinterpret_cast<const uint8_t*>(a.tmpl[k]);
            for (unsigned j = lane; j < (unsigned)msize; j += 32) eq = eq && a.src[sbeg + j] == t[j];
            if (!__all_sync(FULL, eq)) continue;
            const unsigned raw = a.tmpl_raw;
            if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
                uint4* o4 = reinterpret_cast<uint4*>(out);
                for (unsigned q = lane; q < (raw >> 4); q += 32) o4[q] = make_uint4(0, 0, 0, 0);
                for (unsigned q = (raw & ~15u) + lane; q < raw; q += 32) out[q] = 0;
            } else {
                for (unsigned q = lane; q < raw; q += 32) out[q] = 0;
            }
            if (lane == 0) {
                uint32_t* st = a.status + 4 * (size_t)s;
                st[0] = INF_OK; st[1] = raw; st[2] = a.tmpl_chk[k]; st[3] = raw;
            }
            return;
        }
    }
    for (;;) {      // members (gzip files may hold several)
        // ---- container header
        if (a.container == MSL_Z_ZLIB) {
            if (send - bytepos < 6) { err = INF_ERR_HEADER; break; }
            const uint32_t cmf = a.src[bytepos], flg = a.src[bytepos + 1];
            if ((cmf & 15u) != 8 || ((cmf << 8) | flg) % 31u != 0 || (flg & 0x20u)) { err = INF_ERR_HEADER; break; }
            bytepos += 2;
        } else if (a.container == MSL_Z_GZIP) {
            if (send - bytepos < 18 || a.src[bytepos] != 0x1f || a.src[bytepos + 1] != 0x8b || a.src[bytepos + 2] != 8) { err = INF_ERR_HEADER; break; }
            const uint32_t flg = a.src[bytepos + 3];
            unsigned long long p = bytepos + 10;
            if (flg & 4u) { const uint32_t xlen = a.src[p] | ((uint32_t)a.src[p + 1] << 8); p += 2 + xlen; }
            if (flg & 8u) { while (p < send && a.src[p]) ++p; ++p; }
            if (flg & 16u) { while (p < send && a.src[p]) ++p; ++p; }
            if (flg & 2u) p += 2;
            if (p + 8 > send) { err = INF_ERR_HEADER; break; }
            bytepos = p;
        }
        BitReader br;
        br.init(a.src, a.src_bytes, bytepos, T.in_ring, lane);
        bool fixed_ready = false;
        // ---- blocks
        for (;;) {
            br.refill();
            const uint32_t bfinal = br.take(1), btype = br.take(2);
            if (btype == 0) {
                // stored: skip to the byte boundary; LEN / NLEN come out of the bit buffer (no trip to global memory)
                br.drop(br.nb & 7);
                br.refill();
                const uint32_t lw = (uint32_t)br.bb;
                br.bb >>= 32; br.nb -= 32;
                const unsigned long long bp = br.bitpos() >> 3;
                if (bp > send) { err = INF_ERR_INPUT; break; }
                const uint32_t len = lw & 0xffffu, nlen = lw >> 16;
                if ((len ^ nlen) != 0xffffu) { err = INF_ERR_BLOCK; break; }
                if (bp + len > send) { err = INF_ERR_INPUT; break; }
                if (pos + len > cap) { err = INF_ERR_SPACE; break; }
                // an empty stored block (the byte alignment behind a fixed block) costs nothing more: decoding goes on in the bit buffer
                if (len) {
                    // what is still in the ring leaves first; then the block goes straight from the source to the output and its
                    // last kOutRing bytes also into the ring, for later matches.  The first 512 bytes are loaded BEFORE the bit
                    // reader is restarted behind the block, so that both round trips to memory overlap.
                    flush_all();
                    const uint8_t* sp = a.src + bp;
                    if (out_aligned && (pos & 3) == 0) {
                        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(sp) & 3);
                        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(sp - mis);
                        const unsigned long long lim = (a.src_bytes + 3) >> 2;                 // readable words at src
                        const unsigned long long sw0 = (unsigned long long)((sp - mis) - a.src) >> 2;
                        const uint32_t nw = len >> 2;
                        uint32_t* o32 = reinterpret_cast<uint32_t*>(out + pos);
                        uint32_t* r32 = reinterpret_cast<uint32_t*>(ring);
                        const unsigned rw0 = (unsigned)(pos % kOutRing) >> 2;
                        uint32_t lo[4], hi[4];
    #pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t w = lane + 32 * k;
                            lo[k] = (w < nw && sw0 + w < lim) ? __ldg(s32 + w) : 0u;
                            hi[k] = (mis && w < nw && sw0 + w + 1 < lim) ? __ldg(s32 + w + 1) : 0u;
                        }
                        br.init(a.src, a.src_bytes, bp + len, T.in_ring, lane);
    #pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t w = lane + 32 * k;
                            if (w < nw) {
                                const uint32_t v = __funnelshift_r(lo[k], hi[k], 8 * mis);
                                o32[w] = v;
                                if (nw - w < (uint32_t)(kOutRing / 4)) r32[(rw0 + w) % (kOutRing / 4)] = v;
                            }
                        }
                        for (uint32_t base = 128; base < nw; base += 32) {
                            const uint32_t w = base + lane;
                            if (w < nw) {
                                const uint32_t l0 = sw0 + w < lim ? __ldg(s32 + w) : 0u, h0 = (mis && sw0 + w + 1 < lim) ? __ldg(s32 + w + 1) : 0u;
                                const uint32_t v = __funnelshift_r(l0, h0, 8 * mis);
                                o32[w] = v;
                                if (nw - w < (uint32_t)(kOutRing / 4)) r32[(rw0 + w) % (kOutRing / 4)] = v;
                            }
                        }
                        if (const uint32_t i = 4 * nw + lane; i < len) { const uint8_t v = sp[i]; out[pos + i] = v; ring[(pos + i) % kOutRing] = v; }
                    } else {
                        for (uint32_t base = 0; base < len; base += 128) {
                            uint32_t v[4];
    #pragma unroll
                            for (int k = 0; k < 4; ++k) { const uint32_t i = base + lane + 32 * k; v[k] = i < len ? sp[i] : 0u; }
    #pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t i = base + lane + 32 * k;
                                if (i < len) {
                                    out[pos + i] = (uint8_t)v[k];
                                    if (len - i <= (uint32_t)kOutRing) ring[(pos + i) % kOutRing] = (uint8_t)v[k];
                                }
                            }
                        }
                        br.init(a.src, a.src_bytes, bp + len, T.in_ring, lane);
                    }
                    pos += len;
                    flushed = pos;
                    __syncwarp();
                }
            } else if (btype == 3) {
                err = INF_ERR_BLOCK; break;
            } else {
                if (btype == 1) {
                    if (!fixed_ready) {
                        build_fixed_tables(T.lit, T.dist, T.lit_count, T.dist_count, lane);
                        fixed_ready = true;
                    }
                } else {
                    fixed_ready = false;
                    br.refill();
                    const int hlit = (int)br.take(5) + 257, hdist = (int)br.take(5) + 1, hclen = (int)br.take(4) + 4;
                    if (hlit > 286 || hdist > 30) { err = INF_ERR_BLOCK; break; }
                    if (lane < 19) T.lens[lane] = 0;
                    __syncwarp();
                    for (int i = 0; i < hclen; ++i) {
                        br.refill();
                        const uint32_t v = br.take(3);
                        if (lane == 0) T.lens[kClOrder[i]] = (uint8_t)v;
                    }
                    __syncwarp();
                    // the code-length code: 7-bit look-up table in the distance table's storage
                    build_table(T.lens, 19, T.dist, 7, T.dist_count, T.dist_sorted, T.codes, T.next_code, T.offs, lane);
                    int i = 0;
                    const int ntot = hlit + hdist;
                    int prev = 0;
                    bool bad = false;
                    while (i < ntot) {
                        br.refill();
                        const int sym = decode_sym(br, T.dist, 7, T.dist_count, T.dist_sorted);
                        if (sym < 0) { bad = true; break; }
                        if (sym < 16) { if (lane == 0) T.lens[32 + i] = (uint8_t)sym; prev = sym; ++i; }
                        else {
                            int rep, val = 0;
                            if (sym == 16) { if (i == 0) { bad = true; break; } rep = 3 + (int)br.take(2); val = prev; }
                            else if (sym == 17) { rep = 3 + (int)br.take(3); prev = 0; }
                            else { rep = 11 + (int)br.take(7); prev = 0; }
                            if (i + rep > ntot) { bad = true; break; }
                            for (int k = lane; k < rep; k += 32) T.lens[32 + i + k] = (uint8_t)val;
                            i += rep;
                        }
                    }
                    if (bad) { err = INF_ERR_BLOCK; break; }
                    __syncwarp();
                    // lens[32 ..): literal / length lengths, then distance lengths
                    build_table(T.lens + 32, hlit, T.lit, kLitBits, T.lit_count, T.lit_sorted, T.codes, T.next_code, T.offs, lane);
                    build_table(T.lens + 32 + hlit, hdist, T.dist, kDistBits, T.dist_count, T.dist_sorted, T.codes, T.next_code, T.offs, lane);
                }
                // ---- symbols
                for (;;) {
                    br.refill();
                    const int sym = decode_sym(br, T.lit, kLitBits, T.lit_count, T.lit_sorted);
                    if (sym < 256) {
                        if (sym < 0) { err = INF_ERR_CODE; break; }
                        if (pos >= cap) { err = INF_ERR_SPACE; break; }
                        if (lane == 0) ring[pos % kOutRing] = (uint8_t)sym;
                        ++pos;
                        if (pos - flushed >= kFlush) flush_blocks();
                        continue;
                    }
                    if (sym == 256) break;
                    if (sym > 285) { err = INF_ERR_CODE; break; }
                    const int li = sym - 257;
                    const int len = (int)kLenBase[li] + (int)br.take(kLenExtra[li]);
                    br.refill();
                    const int ds = decode_sym(br, T.dist, kDistBits, T.dist_count, T.dist_sorted);
                    if (ds < 0 || ds > 29) { err = INF_ERR_CODE; break; }
                    br.refill();
                    const unsigned dist = (unsigned)kDistBase[ds] + br.take(kDistExtra[ds]);
                    if (pos + len > cap) { err = INF_ERR_SPACE; break; }
                    if (dist > pos) { err = INF_ERR_DIST; break; }
                    __syncwarp();                                   // literals written by lane 0 are visible to every lane
                    const unsigned long long from = pos - dist;
                    if (dist + (unsigned)len <= (unsigned)kRingKeep) {
                        // the whole source lies in the ring
                        if (dist == 1 || dist == 2 || dist == 4) {
                            // a pattern whose period divides four: every aligned word of the destination is the same word
                            uint32_t W = 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) W |= (uint32_t)ring[(from + (((unsigned)j - (unsigned)from) & (dist - 1u))) % kOutRing] << (8 * j);
                            const unsigned head = min((unsigned)len, (4u - ((unsigned)pos & 3u)) & 3u);
                            if ((unsigned)lane < head) ring[(pos + lane) % kOutRing] = (uint8_t)(W >> (8 * (((unsigned)pos + lane) & 3u)));
                            const unsigned long long a0 = pos + head;
                            const unsigned nw = ((unsigned)len - head) >> 2, tail = ((unsigned)len - head) & 3u;
                            uint32_t* r32 = reinterpret_cast<uint32_t*>(ring);
                            const unsigned w0 = (unsigned)(a0 % kOutRing) >> 2;
                            for (unsigned w = lane; w < nw; w += 32) r32[(w0 + w) % (kOutRing / 4)] = W;
                            if ((unsigned)lane < tail) ring[(a0 + 4ull * nw + lane) % kOutRing] = (uint8_t)(W >> (8 * lane));
                        } else if (dist >= (unsigned)len) {
                            for (int i = lane; i < len; i += 32) ring[(pos + i) % kOutRing] = ring[(from + i) % kOutRing];
                        } else {
                            for (int i = lane; i < len; i += 32) ring[(pos + i) % kOutRing] = ring[(from + (unsigned)i % dist) % kOutRing];
                        }
                    } else {
                        // far match: bytes older than the ring's guaranteed window come from global memory (they were flushed:
                        // at most kFlush + 290 bytes are ever unflushed, less than kRingKeep)
                        for (int i = lane; i < len; i += 32) {
                            const unsigned long long idx = from + (dist >= (unsigned)len ? (unsigned)i : (unsigned)i % dist);
                            ring[(pos + i) % kOutRing] = (pos - idx <= (unsigned long long)kRingKeep) ? ring[idx % kOutRing] : out[idx];
                        }
                    }
                    __syncwarp();
                    pos += len;
                    if (pos - flushed >= kFlush) flush_blocks();
                }
                if (err) break;
            }
            if (br.bitpos() > send * 8ull) { err = INF_ERR_INPUT; break; }
            if (bfinal) { bytepos = (br.bitpos() + 7) >> 3; break; }
        }
        if (err) break;
        // ---- trailer
        if (a.container == MSL_Z_ZLIB) {
            if (bytepos + 4 > send) { err = INF_ERR_INPUT; break; }
            stored_chk = ((uint32_t)a.src[bytepos] << 24) | ((uint32_t)a.src[bytepos + 1] << 16) | ((uint32_t)a.src[bytepos + 2] << 8) | a.src[bytepos + 3];
            break;
        }
        if (a.container == MSL_Z_GZIP) {
            if (bytepos + 8 > send) { err = INF_ERR_INPUT; break; }
            stored_chk = a.src[bytepos] | ((uint32_t)a.src[bytepos + 1] << 8) | ((uint32_t)a.src[bytepos + 2] << 16) | ((uint32_t)a.src[bytepos + 3] << 24);
            stored_isize = a.src[bytepos + 4] | ((uint32_t)a.src[bytepos + 5] << 8) | ((uint32_t)a.src[bytepos + 6] << 16) | ((uint32_t)a.src[bytepos + 7] << 24);
            bytepos += 8;
            if (bytepos + 18 <= send && a.src[bytepos] == 0x1f && a.src[bytepos + 1] == 0x8b) continue;    // next member
        }
        break;
    }
    flush_all();
    if (lane == 0) {
        uint32_t* st = a.status + 4 * (size_t)s;
        st[0] = (uint32_t)err; st[1] = (uint32_t)pos; st[2] = stored_chk; st[3] = stored_isize;
    }
}

// ---- PNG scanline filters (PNG 1.2 section 6) undone in place, one warp per image; then channel 0 -> out
struct PngArgs2 {
    uint8_t* raw;                        // inflated scanlines of all images
    const unsigned long long* raw_off;   // [n + 1]
    uint8_t* out;                        // [n][H][W] first channel
    int n, H, W, bpp;                    // bpp = bytes per pixel (1, 2, 3, 4)
    uint32_t* status;                    // [n]: 0 ok, 1 bad filter type / short data
};

__device__ __forceinline__ int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

constexpr int kUnfWarps = 8;

__global__ void __launch_bounds__(kUnfWarps * 32) png_unfilter_kernel(const PngArgs2 a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x;
    uint8_t* raw = a.raw + a.raw_off[s];
    const unsigned long long have = a.raw_off[s + 1] - a.raw_off[s];
    const int rb = a.W * a.bpp, bpp = a.bpp;
    uint8_t* out = a.out + (size_t)s * a.H * a.W;
    if (have < (unsigned long long)a.H * (rb + 1)) { if (threadIdx.x == 0) a.status[s] = 1; return; }
    uint32_t bad = 0;
    if (bpp == 1 && rb <= 256) {
        // None / Sub / Up only (what cv2.imwrite's adaptive filtering picks for masks): None and Sub scanlines do not depend on
        // their neighbours, so the CTA's warps take the rows in turn (Sub = a prefix sum along the row) and write Up rows as
        // the differences they are; a second pass, one thread per column, adds those up going down the image.  Any Average /
        // Paeth scanline sends the image down the serial path below, on warp 0.
        int simple = 1, up_rows = 0;
        for (int y0 = 0; y0 < a.H; y0 += kUnfWarps * 32)           // (warp-uniform trip count: a barrier follows)
            if (const int y = y0 + (int)threadIdx.x; y < a.H) { const int ft = raw[(size_t)y * (rb + 1)]; simple &= ft <= 2 ? 1 : 0; up_rows |= ft == 2 ? 1 : 0; }
        __syncwarp();
        simple = __syncthreads_and(simple);
        up_rows = __syncthreads_or(up_rows);
        if (simple) {
            for (int y = warp; y < a.H; y += kUnfWarps) {
                const uint8_t* row = raw + (size_t)y * (rb + 1);
                const int ft = row[0];
                uint32_t cur[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { const int i = lane + 32 * k; cur[k] = i < rb ? row[1 + i] : 0u; }
                if (ft == 1) {
                    int carry = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (32 * k < rb) {
                            const int v = warp_incl_scan((int)cur[k], lane) + carry;
                            cur[k] = (uint32_t)v & 0xffu;
                            carry = __shfl_sync(FULL, v, 31) & 0xff;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) { const int i = lane + 32 * k; if (i < rb) out[(size_t)y * a.W + i] = (uint8_t)cur[k]; }
            }
            if (up_rows) {
                __syncthreads();
                if (const int x = threadIdx.x; x < rb) {
                    uint32_t prev = 0;
#pragma unroll 4
                    for (int y = 0; y < a.H; ++y) {
                        const int ft = raw[(size_t)y * (rb + 1)];
                        uint32_t v = out[(size_t)y * a.W + x];
                        if (ft == 2) { v = (v + prev) & 0xffu; out[(size_t)y * a.W + x] = (uint8_t)v; }
                        prev = v;
                    }
                }
            }
            if (threadIdx.x == 0) a.status[s] = 0;
            return;
        }
    }
    if (bpp == 1 && rb <= 256) {
        // Average / Paeth scanlines present.  A scanline filtered with None or Sub does not look at the one above it, so the
        // image falls apart into independent chains (such a row + the Up / Average / Paeth rows that follow it); the CTA's
        // warps take the chains in turn.  Inside a chain every lane keeps bytes lane, lane + 32, ... of the current and of the
        // previous (reconstructed) scanline in registers and the next scanline's loads are issued before the current one is
        // processed, so a row costs one memory round trip at most and no row is read twice.
        if (threadIdx.x == 0) a.status[s] = 0;
        __syncthreads();
        auto chain = [&](int yb) {
            uint32_t cur[8], up[8], nxt[8];
            int ft = raw[(size_t)yb * (rb + 1)], ft_next = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int i = lane + 32 * k; up[k] = 0; nxt[k] = 0; cur[k] = i < rb ? raw[(size_t)yb * (rb + 1) + 1 + i] : 0u; }
            for (int y = yb; y < a.H; ++y) {
                bool more = false;
                if (y + 1 < a.H) {
                    const uint8_t* nr = raw + (size_t)(y + 1) * (rb + 1);
                    ft_next = nr[0];
                    more = ft_next >= 2;
                    if (more) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) { const int i = lane + 32 * k; nxt[k] = i < rb ? nr[1 + i] : 0u; }
                    }
                }
                if (ft == 1) {              // Sub: prefix sums mod 256 along the scanline
                    int carry = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (32 * k < rb) {
                            const int v = warp_incl_scan((int)cur[k], lane) + carry;
                            cur[k] = (uint32_t)v & 0xffu;
                            carry = __shfl_sync(FULL, v, 31) & 0xff;
                        }
                    }
                } else if (ft == 2) {       // Up
#pragma unroll
                    for (int k = 0; k < 8; ++k) cur[k] = (cur[k] + up[k]) & 0xffu;
                } else if (ft == 3 || ft == 4) {
                    // Average / Paeth: every byte needs its reconstructed left neighbour - serial along the scanline, the
                    // neighbour handed from lane to lane
                    int left = 0, upleft = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (32 * k < rb) {
                            for (int l = 0; l < 32; ++l) {
                                const int u = (int)up[k];
                                int v = (int)cur[k];
                                if (lane == l) v = (v + (ft == 3 ? ((left + u) >> 1) : paeth(left, u, upleft))) & 0xff;
                                cur[k] = (uint32_t)v;
                                left = __shfl_sync(FULL, v, l);
                                upleft = __shfl_sync(FULL, u, l);
                            }
                        }
                    }
                } else if (ft != 0) bad = 1;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = lane + 32 * k;
                    if (i < rb) out[(size_t)y * a.W + i] = (uint8_t)cur[k];
                    up[k] = cur[k]; cur[k] = nxt[k];
                }
                ft = ft_next;
                if (!more) break;
            }
        };
        int idx = 0;
        for (int y0 = 0; y0 < a.H; y0 += 32) {
            const int y = y0 + lane;
            const bool restart = y < a.H && (y == 0 || raw[(size_t)y * (rb + 1)] <= 1);
            unsigned mask = __ballot_sync(FULL, restart);
            while (mask) {
                const int bsel = __ffs((int)mask) - 1;
                mask &= mask - 1;
                if ((idx++ & (kUnfWarps - 1)) == warp) chain(y0 + bsel);
            }
        }
        if (bad && lane == 0) a.status[s] = 1;
        return;
    }
    if (warp) return;
    for (int y = 0; y < a.H; ++y) {
        uint8_t* cur = raw + (size_t)y * (rb + 1) + 1;
        const uint8_t* up = y ? cur - (rb + 1) : nullptr;
        const int ft = cur[-1];
        if (ft == 1) {
            // Sub: prefix sums (mod 256) along the scanline, one channel at a time, 32 pixels per step with a running carry
            for (int ch = 0; ch < bpp; ++ch) {
                int carry = 0;
                for (int x0 = 0; x0 < a.W; x0 += 32) {
                    const int x = x0 + lane;
                    int v = x < a.W ? cur[x * bpp + ch] : 0;
                    v = warp_incl_scan(v, lane) + carry;
                    if (x < a.W) cur[x * bpp + ch] = (uint8_t)v;
                    carry = __shfl_sync(FULL, v, 31) & 0xff;
                }
            }
        } else if (ft == 2) {
            if (up) for (int i = lane; i < rb; i += 32) cur[i] = (uint8_t)(cur[i] + up[i]);
        } else if (ft == 3 || ft == 4) {
            // Average / Paeth: serial along the scanline (every byte needs its reconstructed left neighbour)
            if (lane == 0) {
                for (int i = 0; i < rb; ++i) {
                    const int l = i >= bpp ? cur[i - bpp] : 0, u = up ? up[i] : 0, ul = (up && i >= bpp) ? up[i - bpp] : 0;
                    cur[i] = (uint8_t)(cur[i] + (ft == 3 ? ((l + u) >> 1) : paeth(l, u, ul)));
                }
            }
        } else if (ft != 0) bad = 1;
        __syncwarp();
        for (int x = lane; x < a.W; x += 32) out[(size_t)y * a.W + x] = cur[x * bpp];
    }
    if (lane == 0) a.status[s] = bad;
}

// ---- NIfTI voxel payload -> float32 / uint8 volume
template <typename T> __device__ __forceinline__ double load_unaligned(const uint8_t* p) {
    T v;
    memcpy(&v, p, sizeof(T));
    return (double)v;
}

// float32 payload, no scaling, aligned buffers, nvox a multiple of four (the volumes of this path): 128-bit loads and stores
__global__ void nifti_convert_f32_kernel(const float4* __restrict__ payload, unsigned long long nquad, float4* __restrict__ out_f32,
                                         uint32_t* __restrict__ out_u8, unsigned long long* inexact) {
    unsigned bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nquad; i += (unsigned long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(payload + i);
        if (out_f32) out_f32[i] = v;
        else {
            const float f[4] = {v.x, v.y, v.z, v.w};
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int q = (f[k] >= 0.0f && f[k] <= 255.0f) ? (int)f[k] : 0;
                w |= (uint32_t)q << (8 * k);
                bad += (float)q != f[k] ? 1u : 0u;
            }
            out_u8[i] = w;
        }
    }
    bad = __reduce_add_sync(FULL, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(inexact, (unsigned long long)bad);
}

__global__ void nifti_convert_kernel(const uint8_t* payload, int datatype, unsigned long long nvox, double slope, double inter, int scaled,
                                     float* out_f32, uint8_t* out_u8, double* out_f64, unsigned long long* inexact) {
    unsigned long long bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nvox; i += (unsigned long long)gridDim.x * blockDim.x) {
        double v;
        switch (datatype) {
            case 2: v = (double)payload[i]; break;
            case 4: v = load_unaligned<int16_t>(payload + 2 * i); break;
            case 8: v = load_unaligned<int32_t>(payload + 4 * i); break;
            case 16: v = load_unaligned<float>(payload + 4 * i); break;
            case 64: v = load_unaligned<double>(payload + 8 * i); break;
            case 256: v = (double)(int8_t)payload[i]; break;
            case 512: v = load_unaligned<uint16_t>(payload + 2 * i); break;
            case 768: v = load_unaligned<uint32_t>(payload + 4 * i); break;
            default: v = 0.0; break;
        }
        if (scaled) v = v * slope + inter;
        if (out_f64) out_f64[i] = v;
        if (out_f32) { const float f = (float)v; out_f32[i] = f; if ((double)f != v && v == v) ++bad; }
        if (out_u8) { const int q = (v >= 0.0 && v <= 255.0) ? (int)v : 0; out_u8[i] = (uint8_t)q; if ((double)q != v) ++bad; }
    }
    bad = __reduce_add_sync(FULL, (unsigned)bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(inexact, bad);
}

}  // namespace

int launch_inflate(const uint8_t* src, size_t src_bytes, const unsigned long long* src_off, int n, int container, uint8_t* dst,
                   const unsigned long long* dst_off, uint32_t* status, cudaStream_t stream) {
    if (n <= 0) return MSL_OK;
    if (reinterpret_cast<uintptr_t>(src) & 3) { set_error("inflate: src must be 4-byte aligned"); return MSL_ERR_ARG; }
    InfArgs a;
    a.src = src; a.src_off = src_off; a.dst = dst; a.dst_off = dst_off; a.status = status; a.n = n; a.container = container;
    a.src_bytes = src_bytes;
    a.tmpl_raw = 0;
    if (container == MSL_Z_GZIP) {
        for (int k = 0; k < 4; ++k) {
            uint32_t meta[4];
            zero_chunk_gzip(k + 1, 16384u, meta, reinterpret_cast<uint8_t*>(a.tmpl[k]));
            a.tmpl_size[k] = meta[0]; a.tmpl_chk[k] = meta[2];
        }
        a.tmpl_raw = 16384u;
    }
    ProfScope prof(K_INFLATE, stream);
    inflate_kernel<<<(n + kInfWarps - 1) / kInfWarps, kInfWarps * 32, 0, stream>>>(a);
    MSL_LAUNCH_CHECK("inflate_kernel");
    return MSL_OK;
}

int launch_png_unfilter(uint8_t* raw, const unsigned long long* raw_off, int n, int H, int W, int bpp, uint8_t* out, uint32_t* status,
                        cudaStream_t stream) {
    if (n <= 0) return MSL_OK;
    PngArgs2 a;
    a.raw = raw; a.raw_off = raw_off; a.out = out; a.n = n; a.H = H; a.W = W; a.bpp = bpp; a.status = status;
    ProfScope prof(K_PNG_UNFILTER, stream);
    png_unfilter_kernel<<<n, kUnfWarps * 32, 0, stream>>>(a);
    MSL_LAUNCH_CHECK("png_unfilter_kernel");
    return MSL_OK;
}

int launch_nifti_convert(const uint8_t* payload, int datatype, unsigned long long nvox, double slope, double inter, int scaled,
                         float* out_f32, uint8_t* out_u8, double* out_f64, unsigned long long* inexact, cudaStream_t stream) {
    if (nvox == 0) return MSL_OK;
    ProfScope prof(K_NIFTI_CONVERT, stream);
    const int blocks = (int)((nvox + 256ull * 8 - 1) / (256ull * 8) < 148 * 16 ? (nvox + 256ull * 8 - 1) / (256ull * 8) : 148 * 16);
    if (datatype == 16 && !scaled && !out_f64 && (nvox & 3) == 0 && (reinterpret_cast<uintptr_t>(payload) & 15) == 0 &&
        (out_f32 ? (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0 : (reinterpret_cast<uintptr_t>(out_u8) & 3) == 0) && (out_f32 != nullptr) != (out_u8 != nullptr)) {
        nifti_convert_f32_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(payload), nvox / 4, reinterpret_cast<float4*>(out_f32),
                                                             reinterpret_cast<uint32_t*>(out_u8), inexact);
        MSL_LAUNCH_CHECK("nifti_convert_f32_kernel");
        return MSL_OK;
    }
    nifti_convert_kernel<<<blocks, 256, 0, stream>>>(payload, datatype, nvox, slope, inter, scaled, out_f32, out_u8, out_f64, inexact);
    MSL_LAUNCH_CHECK("nifti_convert_kernel");
    return MSL_OK;
}

}  // namespace msl
