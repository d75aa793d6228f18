// Byte-stream codec on the device (encode side): RFC 1951 deflate streams inside zlib / gzip / PNG
// containers, so that what crosses PCIe (and what lands on disk) are the compressed files themselves.
//
// Replaces the zlib work behind plt.imsave (scripts/extraer_dataset.py:192,197 -> Pillow PNG encoder), cv2.imwrite
// (utils/utils.py:393, scripts/generar_predicciones.py:153) and nib.save (utils/utils.py:176-177 -> gzip).
//
// deflate_kernel: one CTA per stream (a PNG image, or a 16 KB chunk of a NIfTI file).  Data on this path is either
//   incompressible for a fixed Huffman code (enhanced brain tissue: 8 - 9 bits per literal byte) or one long run
//   (skull-stripped background, masks, padding), so the stream is built from two kinds of blocks decided per WARP:
//   a tile of 4 KB is cut into 256 segments of 16 bytes, one per thread, held in four registers; a byte-SIMD compare gives
//   each thread the mask x[i] == x[i - d] (d = 1: repeated bytes, 4: repeated RGBA pixels / float32 voxels); a ballot tells
//   the warp which of its 32 segments are entirely inside a run.  Every maximal group of neighbouring run segments of a tile
//   becomes ONE fixed-Huffman block of 258-byte matches (at most sixteen of them + end of block); every group of other
//   segments becomes ONE stored block (raw bytes, byte aligned).  Every block is therefore whole bytes (a fixed block is
//   realigned by the header of the stored block behind it, an empty one at the end of the tile), sizes are known from the
//   ballots alone, a warp scan + block scan places every lane, headers are written by the lanes that start a block and every
//   literal lane copies its own 16 bytes.  No literal is ever Huffman-coded - neither here nor in the decoder, which meets
//   a handful of symbols per kilobyte.  The next tile is loaded while this one is encoded (image mode: aligned words,
//   funnel-shifted, the filter byte shifted in), two staging buffers alternate so that a tile costs three barriers;
//   16-byte flushes, the carry to the next tile, Adler-32 (dp4a) / CRC-32 (slicing by four from registers + one GF(2)
//   multiplication per thread) of the raw bytes and the container header / trailer are handled by the same CTA.
// pack_kernel: the variable-length streams are packed back to back (every CTA sums the sizes in front of its streams) into
//   one buffer = one D2H copy; PNG's IDAT CRC-32 (over the compressed bytes) is computed during the copy.
#include <cstring>
#include <mutex>
#include <vector>

#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

namespace {

// CTA barrier for code whose preceding loops may leave a warp split (trip counts that differ inside a warp, e.g.
// `for (q = tid; q < n; q += 256)`): a warp has to be converged when it executes bar.sync - on sm_100a two fragments of one
// warp arriving one after the other were observed to count as two warps and to shift that warp by one barrier phase.
__device__ __forceinline__ void block_sync() { __syncwarp(); __syncthreads(); }

constexpr int kZThreads = 256;
constexpr int kSeg = 16;                         // bytes per thread per tile: the granularity of the run / literal decision
constexpr int kTile = kZThreads * kSeg;          // 4 KB
constexpr int kLook = 16;                        // look-back kept in front of the tile (>= the largest match distance)
constexpr int kYBytes = 16 + kTile + kTile / 2 + 128;   // staging: carry + worst case (every other 16-byte segment its own stored block) + trailer
constexpr uint32_t kCrcPoly = 0xedb88320u;       // reflected CRC-32 polynomial
constexpr int kTmplBytes = kZeroTmplBytes;       // stream of an all-zero chunk of up to four tiles (190 bytes as a gzip member)

// a * b mod P over GF(2), operands in the reflected representation CRC-32 uses (bit 31 = x^0)
__host__ __device__ inline uint32_t gf_mul(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ kCrcPoly : b >> 1;
        if (m == 0) break;
    }
    return p;
}
// x^(8 n) mod P
__host__ __device__ inline uint32_t gf_xpow8(unsigned long long n) {
    uint32_t r = 1u << 31, b = 0x00800000u;      // x^0, x^8
    while (n) { if (n & 1) r = gf_mul(r, b); b = gf_mul(b, b); n >>= 1; }
    return r;
}

struct ZArgs {
    const uint8_t* src;
    size_t src_pitch;          // image mode: bytes between consecutive images; plain mode: bytes between consecutive BODIES
    unsigned long long total;  // plain mode: raw bytes of one source = prefix + body (x 4 when expanding); its last chunk may be short
    unsigned chunk;            // plain mode: raw bytes per stream
    unsigned spv;              // plain mode: streams per source (ceil(total / chunk)); stream s = source s / spv, chunk s % spv
    const uint8_t* prefix;     // plain mode: bytes in front of every body (a file header), or NULL
    size_t prefix_pitch;       // bytes between the prefixes of consecutive sources (0: one prefix for all)
    unsigned prefix_len;
    int expand;                // plain mode: the body is uint8 {0, != 0}; the raw stream holds it as float32 0.0f / 1.0f
    int rows, row_bytes;       // image mode (rows > 0): `rows` scanlines of row_bytes bytes, each prefixed by filter byte 0
    int img_w, img_ch;         // PNG IHDR
    int container;             // MSL_Z_*
    int dist2;                 // match distance 1..4 (0 = 1)
    uint8_t* slots;            // [n][slot_pitch]
    size_t slot_pitch;
    uint32_t* meta;            // [n][4]: container bytes in the slot, raw bytes, checksum of the raw bytes, offset of the IDAT chunk type
    uint32_t crcP[kZThreads];  // x^(8 * kSeg * k): what a thread's 16-byte CRC is multiplied by when k segments follow it in the tile
    uint32_t crc_tilek[4];     // x^(8 * kTile * (k + 1)): k + 1 all-zero tiles
    uint32_t tmpl_meta[4];     // plain mode: the stream of one full all-zero chunk (built on the host): its meta row ([0] = 0: none) ...
    uint32_t tmpl[kTmplBytes / 4];   // ... and its bytes
};

__device__ __forceinline__ unsigned hdr_len_of(int container) {
    return (0x2b180200u >> (8 * (container & 3))) & 0xffu;       // raw 0, zlib 2, gzip 24, PNG 43 (no jump table)
}

__host__ __device__ inline uint32_t brev32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
#endif
}
__host__ __device__ inline int clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
__host__ __device__ inline uint32_t rev_bits(uint32_t v, int n) { return brev32(v) >> (32 - n); }

// fixed Huffman code of a literal / length symbol (RFC 1951 section 3.2.6), already bit-reversed; returns the bit count
__host__ __device__ inline int fixed_litlen(uint32_t sym, uint32_t& code) {
    if (sym < 144) { code = rev_bits(0x30 + sym, 8); return 8; }
    if (sym < 256) { code = rev_bits(0x190 + (sym - 144), 9); return 9; }
    if (sym < 280) { code = rev_bits(sym - 256, 7); return 7; }
    code = rev_bits(0xc0 + (sym - 280), 8); return 8;
}

// One match of `len` bytes (3..258) at distance d (1..4: distance codes 0..3, no extra bits): its bits, LSB first
__host__ __device__ inline int match_bits(int len, int d, uint32_t& pat) {
    uint32_t sym, eb = 0, ev = 0;
    if (len == 258) sym = 285;
    else {
        const uint32_t l = (uint32_t)len - 3;
        if (l < 8) sym = 257 + l;
        else { eb = 29 - clz32(l); sym = 257 + 4 * (eb + 1) + ((l >> eb) & 3); ev = l & ((1u << eb) - 1); }
    }
    uint32_t code;
    const int n = fixed_litlen(sym, code);
    pat = code | (ev << n) | (rev_bits((uint32_t)d - 1, 5) << (n + (int)eb));
    return n + (int)eb + 5;
}

// A run of `bytes` bytes (whole segments of one tile, up to 4 KB) as ONE fixed-Huffman block: header (BFINAL 0, BTYPE 01),
// matches of 258 bytes + the rest (a rest of 1 or 2 bytes is not a match: the last two matches then share 258 + rest bytes),
// end of block.  Returns the bits; W::put(pattern, nbits) receives them in order (a counting writer just sizes the block).
struct BitCounter { __host__ __device__ void put(uint32_t, int) {} };
struct ByteWriter {                              // (host side: the all-zero chunk template)
    uint8_t* b; unsigned long long bit;
    __host__ __device__ void put(uint32_t v, int n) { for (int i = 0; i < n; ++i, ++bit) if ((v >> i) & 1u) b[bit >> 3] |= (uint8_t)(1u << (bit & 7)); }
};
template <class W>
__host__ __device__ inline int emit_run(W& w, int bytes, int d) {
    int q = bytes / 258, r = bytes - 258 * q, tail2 = 0;
    if (r > 0 && r < 3) { q -= 1; tail2 = 3; r = 258 + r - 3; }
    uint32_t pat;
    int n = 3;
    w.put(2u, 3);
    const int m258 = match_bits(258, d, pat);
    for (int i = 0; i < q; ++i) w.put(pat, m258);
    n += q * m258;
    if (r) { const int m = match_bits(r, d, pat); w.put(pat, m); n += m; }
    if (tail2) { const int m = match_bits(tail2, d, pat); w.put(pat, m); n += m; }
    w.put(0u, 7);
    return n + 7;
}

// Tile bytes in shared memory: logical index i (>= -64) lives at i + 4 * ((i + 64) / 64): every thread's 64-byte segment
// starts one bank further (stride 68 bytes), so that the byte-serial tokenisers of a warp do not collide in two banks.
__device__ __forceinline__ int xphys(int i) { return i + (((i + 64) >> 6) << 2); }
#define XP(i) X[xphys(i)]

// ---- a thread's 64-byte segment (16 words in registers): E bit i = x[i] == x[i - d], plus the Adler-32 partial sums
struct SegMasks {
    unsigned long long E;
    uint32_t any;              // OR of the segment's words
    uint32_t s1, s2;           // sum of the bytes, sum of (local index * byte)
    uint32_t w[kSeg / 4];      // the segment itself (bytes beyond len are zero)
    int len;
};

__device__ __forceinline__ uint32_t nibble_of(uint32_t cmp) { return ((cmp & 0x01010101u) * 0x01020408u) >> 24; }

__device__ __forceinline__ SegMasks seg_masks(const uint8_t* X, int beg, int end, long long gfirst, int d) {
    SegMasks m;
    m.E = 0; m.any = 0; m.s1 = 0; m.s2 = 0;
    m.len = end > beg ? end - beg : 0;
#pragma unroll
    for (int k = 0; k < kSeg / 4; ++k) m.w[k] = 0;
    if (m.len == 0) return m;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(X + xphys(beg));
    uint32_t prev = *reinterpret_cast<const uint32_t*>(X + xphys(beg - 4));
#pragma unroll
    for (int k = 0; k < kSeg / 4; ++k) {
        uint32_t word = 4 * k < m.len ? w[k] : 0u;
        if (4 * k + 4 > m.len && 4 * k < m.len) word &= (1u << (8 * (m.len - 4 * k))) - 1u;
        const uint32_t back = d == 4 ? prev : __funnelshift_l(prev, word, 8 * d);        // x[i - d] for the four bytes
        m.E |= (unsigned long long)nibble_of(__vcmpeq4(word, back)) << (4 * k);
        m.any |= word;
        m.w[k] = word;
        m.s1 = __dp4a(word, 0x01010101u, m.s1);
        m.s2 = __dp4a(word, 0x03020100u + 0x04040404u * (uint32_t)k, m.s2);
        prev = word;
    }
    const unsigned long long lenmask = (1ull << m.len) - 1ull;
    m.E &= lenmask;
    if (gfirst + beg < d) m.E &= ~((1ull << (int)(d - (gfirst + beg))) - 1ull);      // no match reaches in front of the stream
    return m;
}

// ---- a tile's 256 segments as blocks, every lane working for itself
// A maximal group of literal segments becomes ONE stored block, a maximal group of run segments ONE fixed-Huffman block (either
// may span the whole 4 KB tile, across warps).  A stored block ends on a byte boundary, so every fixed block STARTS on one (at
// the tile's first byte, behind a stored block, or behind the empty stored block that closes the previous fixed block).  That
// makes every size local: the lane that starts a fixed block emits the block and the header of the stored block behind it (an
// empty one when another fixed block or the end of the tile follows) - whole bytes; a literal lane accounts for its own bytes,
// plus a header if it opens the tile.  A scan of those sizes gives every lane its place; headers are written by their lanes in
// parallel, and every literal lane copies its own 16 bytes.
struct LanePlan {
    bool start, run, hdr;
    int bytes;        // bytes of the block this lane starts
    int next_bytes;   // bytes of the stored block that follows this lane's fixed block (0: an empty one is emitted)
    int hbits;        // bits of the fixed block (header + matches + end of block)
    int size;         // bytes this lane contributes to the tile's output
};

// run segments in a row from segment e of the tile on
__device__ __forceinline__ int run_group_segments(const unsigned* tile_runs, int e, int nseg) {
    int s = e;
    while (s < nseg) {
        const int b = s & 31;
        const unsigned x = tile_runs[s >> 5] >> b;
        const int ones = min(x == (0xffffffffu >> b) ? 32 - b : __ffs((int)~x) - 1, nseg - s);
        s += ones;
        if (ones < 32 - b) break;
    }
    return s - e;
}

// bytes of the literal group that starts at segment e of the tile (0 when that segment is a run or behind the end)
__device__ __forceinline__ int literal_group_bytes(const unsigned* tile_runs, int e, int nseg, unsigned tn) {
    int s = e;
    while (s < nseg) {
        const int b = s & 31;
        const unsigned x = ~tile_runs[s >> 5] >> b;                    // literal segments from s on, in bit order
        const int ones = min(x == (0xffffffffu >> b) ? 32 - b : __ffs((int)~x) - 1, nseg - s);
        s += ones;
        if (ones < 32 - b) break;
    }
    return s > e ? min(kSeg * (s - e), (int)tn - kSeg * e) : 0;
}

__device__ __forceinline__ LanePlan lane_plan(const unsigned* tile_runs, int warp, int lane, int nseg, unsigned tn, int len, int d) {
    LanePlan p;
    p.start = false; p.run = false; p.hdr = false; p.bytes = 0; p.next_bytes = 0; p.hbits = 0; p.size = 0;
    const int seg = warp * 32 + lane;
    if (seg >= nseg) return p;
    const unsigned my = tile_runs[warp];
    p.run = (my >> lane) & 1u;
    if (p.run) {
        const bool prev_run = seg > 0 && (((lane ? my >> (lane - 1) : tile_runs[warp - 1] >> 31)) & 1u);
        p.start = !prev_run;
        if (!p.start) return p;
        const int nl = run_group_segments(tile_runs, seg, nseg);       // (may reach into the following warps)
        p.bytes = kSeg * nl;
        p.next_bytes = literal_group_bytes(tile_runs, seg + nl, nseg, tn);
        BitCounter none;
        p.hbits = emit_run(none, p.bytes, d);
        p.size = ((p.hbits + 3 + 7) >> 3) + 4;
    } else {
        p.hdr = seg == 0;
        p.start = p.hdr;
        if (p.hdr) p.bytes = literal_group_bytes(tile_runs, 0, nseg, tn);
        p.size = len + (p.hdr ? 5 : 0);
    }
    return p;
}

// whole bytes (at most 12) at byte offset `at` of the zeroed staging area
__device__ __forceinline__ void stage_bytes(uint32_t* Y32, unsigned at, unsigned long long lo, uint32_t hi) {
    const unsigned sh = 8u * (at & 3u), w0 = at >> 2;
    uint32_t v0 = (uint32_t)lo, v1 = (uint32_t)(lo >> 32), v2 = hi, v3 = 0;
    if (sh) { v3 = v2 >> (32u - sh); v2 = (v2 << sh) | (v1 >> (32u - sh)); v1 = (v1 << sh) | (v0 >> (32u - sh)); v0 <<= sh; }
    if (v0) atomicOr(&Y32[w0], v0);
    if (v1) atomicOr(&Y32[w0 + 1], v1);
    if (v2) atomicOr(&Y32[w0 + 2], v2);
    if (v3) atomicOr(&Y32[w0 + 3], v3);
}

// emission: at = this lane's place in the staging area (the tile's first byte + the exclusive prefix of the sizes)
// bits into the zeroed staging area from byte `at` on (words are OR-ed in: neighbours may share the first and the last one)
struct StageWriter {
    uint32_t* y;
    unsigned word;
    unsigned long long acc;
    int nacc;
    __device__ __forceinline__ void init(uint32_t* y_, unsigned at) { y = y_; word = at >> 2; nacc = 8 * (int)(at & 3u); acc = 0; }
    __device__ __forceinline__ void put(uint32_t v, int n) {
        acc |= (unsigned long long)v << nacc;
        nacc += n;
        if (nacc >= 32) { if ((uint32_t)acc) atomicOr(&y[word], (uint32_t)acc); ++word; acc >>= 32; nacc -= 32; }
    }
    __device__ __forceinline__ void finish() { if (nacc > 0 && (uint32_t)acc) atomicOr(&y[word], (uint32_t)acc); }
};

__device__ __forceinline__ void lane_emit(const LanePlan& p, const SegMasks& m, int d, uint8_t* Ys, unsigned at) {
    uint32_t* Y32 = reinterpret_cast<uint32_t*>(Ys);
    if (p.start) {
        if (p.run) {
            // the fixed block, the header of the stored block behind it (three zero bits + padding), LEN / NLEN
            StageWriter w;
            w.init(Y32, at);
            emit_run(w, p.bytes, d);
            w.finish();
            const unsigned nb1 = (unsigned)(p.hbits + 3 + 7) >> 3;
            const uint32_t len = (uint32_t)p.next_bytes | ((~(uint32_t)p.next_bytes & 0xffffu) << 16);
            stage_bytes(Y32, at + nb1, (unsigned long long)len, 0u);
        } else {
            const uint32_t len = (uint32_t)p.bytes | ((~(uint32_t)p.bytes & 0xffffu) << 16);
            stage_bytes(Y32, at, (unsigned long long)len << 8, 0u);
        }
    }
    if (!p.run && m.len > 0) {
        const unsigned a = at + (p.hdr ? 5u : 0u);
        const unsigned r = a & 3u, w0 = a >> 2;
        if (r == 0) {
#pragma unroll
            for (int k = 0; k < kSeg / 4; ++k) if (4 * k < m.len) Y32[w0 + k] = m.w[k];
        } else {
            const unsigned sh = 8u * r;
            atomicOr(&Y32[w0], m.w[0] << sh);
#pragma unroll
            for (int k = 1; k < kSeg / 4; ++k) {
                const uint32_t v = (m.w[k] << sh) | (m.w[k - 1] >> (32u - sh));
                if (4 * k < m.len + 4) Y32[w0 + k] = v;              // whole words of this lane (zeros behind a short last segment)
            }
            const uint32_t tail = m.w[kSeg / 4 - 1] >> (32u - sh);
            if (tail) atomicOr(&Y32[w0 + kSeg / 4], tail);
        }
    }
}

// CRC-32 tables for slicing by four, built by 256 threads: T[k][b] = register after byte b and k zero bytes
__device__ __forceinline__ void crc_tables(uint32_t (*T)[256], int tid) {
    uint32_t c = (uint32_t)tid;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
        T[k][tid] = c;
    }
}
__device__ __forceinline__ uint32_t crc_word(const uint32_t (*T)[256], uint32_t c, uint32_t w) {
    c ^= w;
    return T[3][c & 0xffu] ^ T[2][(c >> 8) & 0xffu] ^ T[1][(c >> 16) & 0xffu] ^ T[0][c >> 24];
}
__device__ __forceinline__ uint32_t crc_byte(const uint32_t (*T)[256], uint32_t c, uint32_t b) { return T[0][(c ^ b) & 0xffu] ^ (c >> 8); }

// the 16 raw bytes of segment `tid` of the tile at `tb` (zeros behind the end of the stream).  Image mode inserts the filter
// byte 0 in front of every scanline; plain mode reads prefix + body, the body optionally expanded from uint8 to float32.
__device__ __forceinline__ uint4 load_seg(const ZArgs& a, const uint8_t* __restrict__ src, const uint8_t* __restrict__ pfx, unsigned plen,
                                          bool image, unsigned rl, unsigned rl_magic, unsigned long long r0, unsigned tb, unsigned n, int tid) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    const unsigned b = tb + (unsigned)(kSeg * tid);
    if (b < n) {
        const unsigned cnt = min((unsigned)kSeg, n - b);
        if (image) {
            const unsigned r = __umulhi(b, rl_magic), c0 = b - r * rl;
            const uint8_t* p = src + (size_t)r * (rl - 1) + (c0 ? c0 - 1 : 0);    // the next data byte
            if (rl > (unsigned)kSeg) {
                // at most one filter byte in the segment, at j: 16 data bytes from aligned words, then a zero byte shifted in
                const unsigned j = c0 == 0 ? 0u : rl - c0;
                const unsigned ndata = cnt - (j < cnt ? 1u : 0u);
                const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(p) & 3);
                const uint32_t* a32 = reinterpret_cast<const uint32_t*>(p - mis);
                uint32_t q[kSeg / 4 + 1];
#pragma unroll
                for (int k = 0; k <= kSeg / 4; ++k) q[k] = (unsigned)(4 * k) < mis + ndata ? __ldg(a32 + k) : 0u;     // only words that hold a data byte
                uint32_t D[kSeg / 4];
#pragma unroll
                for (int k = 0; k < kSeg / 4; ++k) D[k] = __funnelshift_r(q[k], q[k + 1], 8 * mis);
                const unsigned jw = j >> 2, jb = j & 3u;
                const uint32_t below = (1u << (8 * jb)) - 1u, upto = 0xffffffffu >> (8 * (3 - jb));
#pragma unroll
                for (int k = 0; k < kSeg / 4; ++k) {
                    const uint32_t S = k ? __funnelshift_l(D[k - 1], D[k], 8) : D[0] << 8;
                    w[k] = (unsigned)k < jw ? D[k] : ((unsigned)k > jw ? S : ((D[k] & below) | (S & ~upto)));
                }
                if (cnt < (unsigned)kSeg) {
#pragma unroll
                    for (int k = 0; k < kSeg / 4; ++k) {
                        const int left = (int)cnt - 4 * k;
                        if (left <= 0) w[k] = 0; else if (left < 4) w[k] &= (1u << (8 * left)) - 1u;
                    }
                }
            } else {
                unsigned c = c0;
#pragma unroll
                for (int k = 0; k < kSeg; ++k) {
                    const bool data = c != 0;
                    uint32_t v = 0;
                    if ((unsigned)k < cnt && data) v = __ldg(p);
                    p += data ? 1 : 0;
                    c = c + 1 == rl ? 0u : c + 1;
                    w[k >> 2] |= v << (8 * (k & 3));
                }
            }
        } else {
            const unsigned long long g = r0 + b;
            if (!a.expand && g >= plen && cnt == kSeg && (reinterpret_cast<uintptr_t>(src + (g - plen)) & 15) == 0) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (g - plen)));
                return v;
            } else if (a.expand && g >= plen && ((g - plen) & 3) == 0 && (cnt & 3) == 0) {
                // uint8 mask stored as float32: one mask byte per output word (0.0f / 1.0f)
                const uint8_t* body = src + ((g - plen) >> 2);
                uint32_t mk = 0;
                if (cnt == kSeg && (reinterpret_cast<uintptr_t>(body) & 3) == 0) mk = __ldg(reinterpret_cast<const uint32_t*>(body));
                else {
#pragma unroll
                    for (int k = 0; k < kSeg / 4; ++k) if ((unsigned)(4 * k) < cnt) mk |= (uint32_t)__ldg(body + k) << (8 * k);
                }
#pragma unroll
                for (int k = 0; k < kSeg / 4; ++k) w[k] = ((mk >> (8 * k)) & 0xffu) ? 0x3f800000u : 0u;
            } else {
#pragma unroll
                for (int k = 0; k < kSeg; ++k) if ((unsigned)k < cnt) {
                    const unsigned long long gg = g + k;
                    uint32_t v;
                    if (gg < plen) v = __ldg(pfx + gg);
                    else if (!a.expand) v = __ldg(src + (gg - plen));
                    else {
                        const unsigned long long bb = gg - plen;
                        v = __ldg(src + (bb >> 2)) ? (0x3f800000u >> (8 * (unsigned)(bb & 3))) & 0xffu : 0u;
                    }
                    w[k >> 2] |= v << (8 * (k & 3));
                }
            }
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}


__global__ void __launch_bounds__(kZThreads, 5) deflate_kernel(const ZArgs a) {
    __shared__ __align__(16) uint8_t Xs[kLook + kTile + 4 * (kZThreads + 2) + 16];
    __shared__ __align__(16) uint8_t Ys2[2][kYBytes + 16];
    __shared__ uint32_t crc_table[4][256];            // slicing by four: table k = a byte followed by k zero bytes
    __shared__ uint32_t crc_part[kZThreads / 32];
    __shared__ unsigned tile_runs[kZThreads / 32];     // bit l of word w: segment 32 w + l of the tile lies inside a run
    __shared__ int scan_w[kZThreads / 32];
    __shared__ unsigned long long red_a[kZThreads / 32], red_b[kZThreads / 32];
    __shared__ unsigned s_state[4];     // [3] bytes of the last flush
    uint8_t* X = Xs + kLook;
    uint8_t* Ys = Ys2[0];
    uint32_t* Y32 = reinterpret_cast<uint32_t*>(Ys);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned s = blockIdx.x;
    const bool image = a.rows > 0;
    const unsigned rl = image ? (unsigned)a.row_bytes + 1u : 0u;
    unsigned n;                                            // raw bytes of this stream
    if (image) n = (unsigned)a.rows * rl;
    unsigned long long r0 = 0;                            // plain mode: raw offset of this stream inside its source
    unsigned vsrc = s;
    if (!image) {
        vsrc = s / a.spv;
        r0 = (unsigned long long)(s - vsrc * a.spv) * a.chunk;
        n = r0 >= a.total ? 0u : (unsigned)min((unsigned long long)a.chunk, a.total - r0);
    }
    const uint8_t* src = a.src + (size_t)vsrc * a.src_pitch;
    const uint8_t* pfx = a.prefix ? a.prefix + (size_t)vsrc * a.prefix_pitch : nullptr;
    const unsigned plen = a.prefix ? a.prefix_len : 0u;
    uint8_t* slot = a.slots + (size_t)s * a.slot_pitch;
    const unsigned hdr = hdr_len_of(a.container);
    const bool want_crc = a.container == MSL_Z_GZIP, want_adler = a.container == MSL_Z_ZLIB || a.container == MSL_Z_PNG;
    const unsigned rl_magic = image ? (unsigned)(0x100000000ull / rl) + 1u : 0u;   // g / rl == umulhi(g, magic) for g * rl < 2^32

    const int dist = a.dist2 ? a.dist2 : 1;
    if (want_crc) crc_tables(crc_table, tid);
    for (int q0_ = 0; q0_ < ((kYBytes + 16) / 4); q0_ += kZThreads) if (const int q = q0_ + (int)tid; q < ((kYBytes + 16) / 4)) Y32[q] = 0;
    if (tid < kLook / 4) reinterpret_cast<uint32_t*>(Xs)[tid] = 0;
    block_sync();
    if (tid == 0) {
        // container header (the sizes inside it are patched at the end) and the block header: BFINAL = 1, BTYPE = 01
        if (a.container == MSL_Z_ZLIB) { Ys[0] = 0x78; Ys[1] = 0x01; }
        else if (a.container == MSL_Z_GZIP) {
            const uint8_t h[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 12, 0, 'M', 'S', 8, 0};
            for (int i = 0; i < 16; ++i) Ys[i] = h[i];
        } else if (a.container == MSL_Z_PNG) {
            const uint8_t sig[16] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a, 0, 0, 0, 13, 'I', 'H', 'D', 'R'};
            for (int i = 0; i < 16; ++i) Ys[i] = sig[i];
            const uint32_t w = (uint32_t)a.img_w, h = (uint32_t)a.rows;
            Ys[16] = w >> 24; Ys[17] = w >> 16; Ys[18] = w >> 8; Ys[19] = w;
            Ys[20] = h >> 24; Ys[21] = h >> 16; Ys[22] = h >> 8; Ys[23] = h;
            Ys[24] = 8; Ys[25] = (uint8_t)((0x06020400u >> (8 * ((a.img_ch - 1) & 3))) & 0xffu);     // colour type: gray 0, gray+alpha 4, RGB 2, RGBA 6 Ys[26] = 0; Ys[27] = 0; Ys[28] = 0;
            uint32_t c = 0xffffffffu;
            for (int i = 12; i < 29; ++i) { c ^= Ys[i]; for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1; }
            c = ~c;
            Ys[29] = c >> 24; Ys[30] = c >> 16; Ys[31] = c >> 8; Ys[32] = c;
            Ys[37] = 'I'; Ys[38] = 'D'; Ys[39] = 'A'; Ys[40] = 'T';       // [33, 37): IDAT length, patched by pack_kernel
            Ys[41] = 0x78; Ys[42] = 0x01;
        }
    }
    // bits carried in the staging area and bytes already in the slot: every thread keeps its own copy
    unsigned bit0 = 8 * hdr, done = 0;
    uint32_t crc_run = 0;                               // thread 0: running CRC-32 (zero-register form) and all-zero tiles not yet applied
    int crc_pend = 0;
    uint4 P = load_seg(a, src, pfx, plen, image, rl, rl_magic, r0, 0u, n, tid);
    if (a.tmpl_meta[0] && !image && n == a.chunk && r0 >= plen) {
        // A full chunk without header bytes: when every byte of it is zero (two thirds of a skull-stripped volume, nearly all of a
        // mask) its stream is the same for every such chunk - built once on the host (zero_chunk_stream), copied here
        uint32_t nz = P.x | P.y | P.z | P.w;
        for (unsigned tb = kTile; tb < n; tb += kTile) {
            const uint4 q = load_seg(a, src, pfx, plen, image, rl, rl_magic, r0, tb, n, tid);
            nz |= q.x | q.y | q.z | q.w;
        }
        __syncwarp();
        if (!__syncthreads_or(nz != 0)) {
            const unsigned fsize = a.tmpl_meta[0];
            uint32_t* d32 = reinterpret_cast<uint32_t*>(slot);
            if (tid < (int)((fsize + 3) >> 2)) d32[tid] = a.tmpl[tid];
            if (tid < 4) a.meta[4 * (size_t)s + tid] = a.tmpl_meta[tid];
            return;
        }
    }
    block_sync();

    unsigned long long sa = 0, sb = 0;                  // Adler-32 partial sums
    int cur = 0;
    for (unsigned tb = 0; tb < n || tb == 0; tb += kTile) {
        const unsigned tn = min((unsigned)kTile, n - tb);
        Ys = Ys2[cur];
        Y32 = reinterpret_cast<uint32_t*>(Ys);
        // ---- the tile into shared memory (loaded one iteration ahead); the other staging buffer is cleared meanwhile
        {
            uint32_t* x32 = reinterpret_cast<uint32_t*>(X + xphys(tid * kSeg));
            x32[0] = P.x; x32[1] = P.y; x32[2] = P.z; x32[3] = P.w;
        }
        block_sync();
        const bool last = tb + kTile >= n;
        if (!last) P = load_seg(a, src, pfx, plen, image, rl, rl_magic, r0, tb + kTile, n, tid);
        {
            uint4* o4 = reinterpret_cast<uint4*>(Ys2[cur ^ 1]);
            for (int q0_ = 0; q0_ < (kYBytes + 16) / 16; q0_ += kZThreads) if (const int q = q0_ + tid; q < (kYBytes + 16) / 16) o4[q] = make_uint4(0, 0, 0, 0);
        }
        // ---- the thread's segment as masks; checksums of the raw bytes
        const int beg = tid * kSeg, end = min(beg + kSeg, (int)tn);
        const SegMasks m = seg_masks(X, beg, end, (long long)tb, dist);
        if (want_adler) { sa += m.s1; sb += (unsigned long long)(n - (tb + beg)) * m.s1 - m.s2; }
        const bool isrun = m.len == kSeg && m.E == ((1ull << kSeg) - 1ull);
        const unsigned run_mask = __ballot_sync(FULL, isrun);
        if (lane == 0) tile_runs[warp] = run_mask;
        bool crc_fold = false;
        if (!(want_crc && tn > 0)) block_sync();
        else {
            if (__syncwarp(), !__syncthreads_or(m.any != 0 || tb < 4u || tn != (unsigned)kTile)) {
                // an all-zero tile: a zero-initialised CRC register stays zero, only the running value moves on - lazily, by
                // up to four tiles at once
                if (tid == 0 && ++crc_pend == 4) { crc_run = gf_mul(crc_run, a.crc_tilek[3]); crc_pend = 0; }
            } else {
                // the thread's own 16 bytes (the all-ones initial register of CRC-32 is folded in by complementing the first
                // four bytes of the stream; thread 0 starts from the running value instead of zero), moved to the end of the
                // tile by ONE multiplication with x^(8 * bytes behind the segment); the XOR of the 256 values is the new
                // running value (partials travel with the block scan's barrier)
                uint32_t c = 0;
                if (tid == 0) {
                    if (crc_pend) { crc_run = gf_mul(crc_run, a.crc_tilek[crc_pend - 1]); crc_pend = 0; }
                    c = crc_run;
                }
#pragma unroll
                for (int k = 0; k < kSeg / 4; ++k) if (4 * k + 4 <= m.len) c = crc_word(crc_table, c, (tb == 0 && tid == 0 && k == 0) ? ~m.w[0] : m.w[k]);
                for (int i = m.len & ~3; i < m.len; ++i) c = crc_byte(crc_table, c, XP(beg + i));   // (streams of fewer than four bytes are redone below)
                if (m.len == kSeg) {
                    const int full = (int)(tn >> 4), rem = (int)(tn & 15u);        // bytes behind me: 16 * (full - 1 - tid) + rem
                    for (int i = 0; i < rem; ++i) c = crc_byte(crc_table, c, 0u);
                    c = gf_mul(a.crcP[full - 1 - tid], c);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(FULL, c, o);
                if (lane == 0) crc_part[warp] = c;
                crc_fold = true;
            }
        }
        // ---- blocks: sizes from the run masks of the tile, a scan over lanes and warps, emission
        const LanePlan plan = lane_plan(tile_runs, warp, lane, (int)((tn + kSeg - 1) / kSeg), tn, m.len, dist);
        const int lincl = warp_incl_scan(plan.size, lane);
        if (lane == 31) scan_w[warp] = lincl;
        block_sync();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kZThreads / 32; ++w) { const int v = scan_w[w]; if (w < warp) wbase += v; total += v; }
        if (crc_fold && tid == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < kZThreads / 32; ++w) t ^= crc_part[w];
            crc_run = t;
        }
        lane_emit(plan, m, dist, Ys, (bit0 >> 3) + (unsigned)(wbase + lincl - plan.size));
        total *= 8;
        block_sync();
        const unsigned nbits = bit0 + (unsigned)total;
        if (last) {
            // end of block (symbol 256: seven zero bits), byte alignment, trailer
            if (want_adler) {
                for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(FULL, sa, o); sb += __shfl_xor_sync(FULL, sb, o); }
                if (lane == 0) { red_a[warp] = sa; red_b[warp] = sb; }
            }
            block_sync();
            if (tid == 0) {
                unsigned nb = nbits >> 3;                               // every warp's output is whole bytes
                Ys[nb] = 1; Ys[nb + 1] = 0; Ys[nb + 2] = 0; Ys[nb + 3] = 0xff; Ys[nb + 4] = 0xff;   // BFINAL = 1: empty stored block
                nb += 5;
                uint32_t chk = 0;
                if (want_adler) {
                    unsigned long long ta = 1, tbb = n;
                    for (int w = 0; w < kZThreads / 32; ++w) { ta += red_a[w]; tbb += red_b[w]; }
                    chk = (uint32_t)((tbb % 65521ull) << 16) | (uint32_t)(ta % 65521ull);
                    Ys[nb] = chk >> 24; Ys[nb + 1] = chk >> 16; Ys[nb + 2] = chk >> 8; Ys[nb + 3] = chk;
                    nb += 4;
                } else if (want_crc) {
                    if (crc_pend) crc_run = gf_mul(crc_run, a.crc_tilek[crc_pend - 1]);
                    if (n >= 4) chk = crc_run ^ 0xffffffffu;
                    else {                                  // streams shorter than the register: the plain definition (the tile is still in X)
                        uint32_t c = 0xffffffffu;
                        for (unsigned i = 0; i < n; ++i) c = crc_byte(crc_table, c, XP((int)i));
                        chk = ~c;
                    }
                    Ys[nb] = chk; Ys[nb + 1] = chk >> 8; Ys[nb + 2] = chk >> 16; Ys[nb + 3] = chk >> 24;
                    Ys[nb + 4] = n; Ys[nb + 5] = n >> 8; Ys[nb + 6] = n >> 16; Ys[nb + 7] = n >> 24;
                    nb += 8;
                }
                unsigned idat_type = 0;
                if (a.container == MSL_Z_PNG) {
                    // [nb, nb + 4): IDAT CRC (pack_kernel), then IEND
                    const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xae, 0x42, 0x60, 0x82};
                    for (int i = 0; i < 12; ++i) Ys[nb + 4 + i] = iend[i];
                    nb += 16;
                    idat_type = 37;
                }
                const unsigned fsize = done + nb;
                uint32_t* m = a.meta + 4 * (size_t)s;
                m[0] = fsize; m[1] = n; m[2] = chk; m[3] = idat_type;
                if (a.container == MSL_Z_GZIP) {
                    // compressed member size and raw size into the 'MS' extra subfield (bytes 16..23 of the member): still in
                    // the staging area for a one-tile stream, already in the slot (flushed before earlier barriers) otherwise
                    uint8_t* h = done == 0 ? Ys + 16 : slot + 16;
                    h[0] = fsize; h[1] = fsize >> 8; h[2] = fsize >> 16; h[3] = fsize >> 24;
                    h[4] = n; h[5] = n >> 8; h[6] = n >> 16; h[7] = n >> 24;
                }
                s_state[3] = nb;
            }
            block_sync();
            const unsigned nb = s_state[3];
            uint4* d4 = reinterpret_cast<uint4*>(slot + done);
            const uint4* y4 = reinterpret_cast<const uint4*>(Ys);
            for (unsigned q0_ = 0; q0_ < (((nb + 15) >> 4)); q0_ += kZThreads) if (const unsigned q = q0_ + (unsigned)tid; q < (((nb + 15) >> 4))) d4[q] = y4[q];
            break;
        }
        // ---- flush whole 16-byte groups, carry the rest to the front of the other staging buffer
        const unsigned nfl = (nbits >> 3) & ~15u;
        {
            uint4* d4 = reinterpret_cast<uint4*>(slot + done);
            const uint4* y4 = reinterpret_cast<const uint4*>(Ys);
            for (unsigned q0_ = 0; q0_ < ((nfl >> 4)); q0_ += kZThreads) if (const unsigned q = q0_ + (unsigned)tid; q < ((nfl >> 4))) d4[q] = y4[q];
        }
        if (tid < 2) *reinterpret_cast<uint4*>(Ys2[cur ^ 1] + 16 * tid) = *reinterpret_cast<const uint4*>(Ys + nfl + 16 * tid);   // < 16 bytes and a partial byte remain
        // the look-back of the next tile = the last segment of this one (still in its owner's registers)
        if (tid == kZThreads - 1) *reinterpret_cast<uint4*>(Xs) = make_uint4(m.w[0], m.w[1], m.w[2], m.w[3]);
        bit0 = nbits - 8 * nfl;
        done += nfl;
        cur ^= 1;
    }
}

// copies stream s from its slot to out + off[s]; PNG: IDAT length and CRC-32 over (type + compressed data)
struct PackArgs {
    const uint8_t* slots;
    size_t slot_pitch;
    const uint32_t* meta;
    unsigned long long* off;   // [n + 1]: written here (exclusive scan of the sizes)
    uint8_t* out;
    unsigned long long out_cap;
    uint32_t P64[kZThreads];   // x^(8 * 64 * k)
    uint32_t x16k;             // x^(8 * 16384)
    int n;
};

// one stream: slot -> out + off[s]
__device__ __forceinline__ void pack_stream(const PackArgs& a, unsigned s, unsigned long long o, uint32_t (*crc_table)[256], uint32_t* part,
                                            uint32_t& s_run) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* slot = a.slots + (size_t)s * a.slot_pitch;
    const uint32_t fsize = a.meta[4 * (size_t)s], idat = a.meta[4 * (size_t)s + 3];
    if (o + fsize > a.out_cap) return;                      // the caller compares off[n] with the capacity
    uint8_t* dst = a.out + o;
    uint32_t clen = 0;                                      // PNG: bytes the IDAT CRC covers = "IDAT" + zlib stream
    if (idat) {
        clen = fsize - idat - 16;                           // file = ... [idat - 4: length][idat: "IDAT" + data][CRC][IEND chunk: 12]
        if (tid == 0) s_run = 0;
    }
    // destination words; the source (16-byte aligned slots) is read as aligned words and shifted into place
    const unsigned head = (unsigned)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
    const unsigned h = min(head, fsize);
    if (tid < (int)h) dst[tid] = slot[tid];
    const unsigned nwords = (fsize - h) >> 2;
    {
        const unsigned ms = (unsigned)((reinterpret_cast<uintptr_t>(slot) + h) & 3);
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(slot + h - ms);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + h);
        if (ms == 0) {
            for (unsigned q0_ = 0; q0_ < nwords; q0_ += kZThreads) if (const unsigned q = q0_ + (unsigned)tid; q < nwords) d32[q] = s32[q];
        } else {
            for (unsigned q0_ = 0; q0_ < nwords; q0_ += kZThreads) if (const unsigned q = q0_ + (unsigned)tid; q < nwords)
                d32[q] = __funnelshift_r(s32[q], s32[q + 1], 8 * ms);      // (the slot has slack behind the stream)
        }
    }
    if (const unsigned i = h + 4 * nwords + tid; i < fsize) dst[i] = slot[i];                       // (< 4 bytes)
    if (!idat) return;
    block_sync();
    // CRC-32 of the IDAT chunk in tiles of 16 KB that are aligned to the END of the chunk (only the first tile is short, and a
    // zero register does not see missing leading bytes): a thread takes 64 bytes (aligned words shifted into place, four
    // bytes per table step), moves its value to the end of the tile with one multiplication, the tile's XOR is folded into
    // the running value.  The all-ones initial register = the first four bytes complemented.
    const uint8_t* C = slot + idat;
    const int ntile = (int)((clen + 16383u) >> 14);
    for (int j = 0; j < ntile; ++j) {
        const long long te = (long long)clen - (long long)(ntile - 1 - j) * 16384;
        const long long e = te - (long long)(kZThreads - 1 - tid) * 64, b = e - 64;
        const long long lo = max(b, max(te - 16384, 0ll));
        const int count = e > lo ? (int)(e - lo) : 0;
        uint32_t c = 0;
        if (count > 0) {
            const uint8_t* p = C + lo;
            const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(p) & 3);
            const uint32_t* a32 = reinterpret_cast<const uint32_t*>(p - mis);
            const uint32_t flip0 = lo >= 4 ? 0u : (lo == 0 ? 0xffffffffu : (1u << (8 * (4 - (int)lo))) - 1u);
            if (count == 64) {
                uint32_t q[17];
#pragma unroll
                for (int k = 0; k < 17; ++k) q[k] = (k < 16 || mis) ? a32[k] : 0u;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    uint32_t w = __funnelshift_r(q[k], q[k + 1], 8 * mis);
                    if (k == 0) w ^= flip0;
                    c = crc_word(crc_table, c, w);
                }
            } else {
                const int nw = count >> 2;
                uint32_t prev = a32[0];
                for (int k = 0; k <= nw; ++k) {
                    const uint32_t nxt = (unsigned)(4 * (k + 1)) < mis + (unsigned)count ? a32[k + 1] : 0u;
                    uint32_t w = __funnelshift_r(prev, nxt, 8 * mis);
                    if (k == 0) w ^= flip0;
                    if (k < nw) c = crc_word(crc_table, c, w);
                    else for (int t = 0; t < (count & 3); ++t) c = crc_byte(crc_table, c, (w >> (8 * t)) & 0xffu);
                    prev = nxt;
                }
            }
            c = gf_mul(a.P64[kZThreads - 1 - tid], c);
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) c ^= __shfl_xor_sync(FULL, c, o2);
        if (lane == 0) part[warp] = c;
        block_sync();
        if (tid == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kZThreads / 32; ++w) t ^= part[w];
            s_run = (j ? gf_mul(s_run, a.x16k) : 0u) ^ t;
        }
        block_sync();
    }
    if (tid == 0) {
        const uint32_t crc = s_run ^ 0xffffffffu;
        uint8_t* q = dst + idat + clen;
        q[0] = crc >> 24; q[1] = crc >> 16; q[2] = crc >> 8; q[3] = crc;
        const uint32_t dl = clen - 4;
        uint8_t* p = dst + idat - 4;
        p[0] = dl >> 24; p[1] = dl >> 16; p[2] = dl >> 8; p[3] = dl;
    }
}

// Grid-stride over the streams: a bounded number of CTAs, so that a launch whose destination is mapped HOST memory (the copy
// then runs at the speed of the host link) leaves most of every SM to the kernels of other CUDA streams.  Every CTA sums the
// sizes in front of its streams itself (a few thousand 32-bit loads at most), which saves the scan kernel between deflate and
// pack - a one-CTA launch that used to queue behind the wide kernels of the other streams.
__global__ void __launch_bounds__(kZThreads) pack_kernel(const PackArgs a) {
    __shared__ uint32_t crc_table[4][256];
    __shared__ uint32_t part[kZThreads / 32];
    __shared__ uint32_t s_run;
    __shared__ unsigned long long red[kZThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto block_sum = [&](unsigned lo, unsigned hi) -> unsigned long long {          // sum of the sizes of streams [lo, hi)
        unsigned long long v = 0;
        for (unsigned j0 = lo; j0 < hi; j0 += kZThreads) if (const unsigned j = j0 + (unsigned)tid; j < hi) v += a.meta[4 * (size_t)j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        block_sync();                                                                 // (red may still be read from the last call)
        if (lane == 0) red[warp] = v;
        block_sync();
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kZThreads / 32; ++w) t += red[w];
        return t;
    };
    crc_tables(crc_table, tid);
    unsigned long long o = block_sum(0, blockIdx.x);
    for (unsigned s = blockIdx.x; s < (unsigned)a.n; s += gridDim.x) {
        if (tid == 0) {
            a.off[s] = o;
            if (s + 1 == (unsigned)a.n) a.off[s + 1] = o + a.meta[4 * (size_t)s];
        }
        pack_stream(a, s, o, crc_table, part, s_run);
        if (s + gridDim.x < (unsigned)a.n) o += block_sum(s, s + gridDim.x);
        else block_sync();
    }
}

// The stream deflate_kernel produces for a chunk of n zero bytes (n = 1..4 whole tiles), restated on the host: tile 0 opens with
// a stored block of the first segment (nothing to refer back to), then every warp's run of segments is one fixed block + the
// empty stored block that realigns it.  Cached per (container, distance, n).
void zero_chunk_stream(int container, int d, unsigned n, uint32_t meta[4], uint8_t* out) {
    struct Entry { int container, d; unsigned n; uint32_t meta[4]; uint8_t bytes[kTmplBytes]; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    std::lock_guard<std::mutex> lk(mu);
    for (const Entry& e : cache)
        if (e.container == container && e.d == d && e.n == n) { memcpy(meta, e.meta, sizeof(e.meta)); memcpy(out, e.bytes, kTmplBytes); return; }
    Entry e;
    memset(&e, 0, sizeof(e));
    e.container = container; e.d = d; e.n = n;
    uint8_t* b = e.bytes;
    unsigned p = 0;
    if (container == MSL_Z_ZLIB) { b[0] = 0x78; b[1] = 0x01; p = 2; }
    else if (container == MSL_Z_GZIP) {
        const uint8_t h[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 12, 0, 'M', 'S', 8, 0};
        memcpy(b, h, 16); p = 24;
    }
    auto put_run = [&](int nl) {               // a run of nl segments: fixed block, empty stored block behind it
        ByteWriter w{b, 8ull * p};
        const int hbits = emit_run(w, kSeg * nl, d);
        p += (unsigned)(hbits + 3 + 7) >> 3;
        b[p++] = 0; b[p++] = 0; b[p++] = 0xff; b[p++] = 0xff;
    };
    for (unsigned t = 0; t < n / kTile; ++t) {
        if (t == 0) {
            b[p++] = 0; b[p++] = kSeg; b[p++] = 0; b[p++] = (uint8_t)~kSeg; b[p++] = 0xff;
            p += kSeg;                          // the sixteen zero bytes themselves
            put_run(kZThreads - 1);
        } else put_run(kZThreads);
    }
    b[p++] = 1; b[p++] = 0; b[p++] = 0; b[p++] = 0xff; b[p++] = 0xff;
    uint32_t chk = 0;
    if (container == MSL_Z_ZLIB) {
        chk = ((n % 65521u) << 16) | 1u;
        b[p++] = chk >> 24; b[p++] = chk >> 16; b[p++] = chk >> 8; b[p++] = chk;
    } else if (container == MSL_Z_GZIP) {
        uint32_t c = 0xffffffffu;
        for (unsigned i = 0; i < n; ++i) for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
        chk = ~c;
        for (int i = 0; i < 4; ++i) b[p++] = (uint8_t)(chk >> (8 * i));
        for (int i = 0; i < 4; ++i) b[p++] = (uint8_t)(n >> (8 * i));
        for (int i = 0; i < 4; ++i) { b[16 + i] = (uint8_t)(p >> (8 * i)); b[20 + i] = (uint8_t)(n >> (8 * i)); }
    }
    e.meta[0] = p; e.meta[1] = n; e.meta[2] = chk; e.meta[3] = 0;
    cache.push_back(e);
    memcpy(meta, e.meta, sizeof(e.meta)); memcpy(out, e.bytes, kTmplBytes);
}

}  // namespace

void zero_chunk_gzip(int d, unsigned n, uint32_t meta[4], uint8_t* out) { zero_chunk_stream(MSL_Z_GZIP, d, n, meta, out); }

namespace {

inline size_t raw_len_of(size_t chunk, int rows, int row_bytes) { return rows > 0 ? (size_t)rows * ((size_t)row_bytes + 1) : chunk; }

}  // namespace

size_t deflate_slot_bytes(int container, size_t raw) {
    const size_t hdr = container == MSL_Z_ZLIB ? 2 : container == MSL_Z_GZIP ? 24 : container == MSL_Z_PNG ? 43 : 0;
    return ((hdr + (raw * 9 + 10 + 7) / 8 + 32 + 16) + 15) & ~(size_t)15;
}

size_t deflate_workspace_bytes(int n, int container, size_t raw) {
    // meta rows, the slots, and the template of an all-zero chunk (meta row + one slot)
    return (size_t)(n + 1) * deflate_slot_bytes(container, raw) + (((size_t)n * 16 + 255) & ~(size_t)255) + 16;
}

int launch_deflate_pack(const uint8_t* src, int n, size_t src_pitch, size_t chunk, size_t total, int rows, int row_bytes,
                        int img_w, int img_ch, int container, int dist2, const uint8_t* prefix, size_t prefix_pitch, size_t prefix_len,
                        int expand, uint8_t* out, size_t out_cap, unsigned long long* out_off,
                        uint32_t* out_meta, void* ws, size_t ws_bytes, cudaStream_t stream) {
    const size_t raw = raw_len_of(chunk, rows, row_bytes);
    if (raw >= (1u << 24)) { set_error("deflate: streams of up to 16 MB (got %zu bytes)", raw); return MSL_ERR_UNSUPPORTED; }
    if (rows > 0 && (unsigned long long)raw * ((unsigned long long)row_bytes + 1) >= 0x100000000ull) {
        set_error("deflate: image of %d x %d bytes too large", rows, row_bytes); return MSL_ERR_UNSUPPORTED;
    }
    if (dist2 < 0 || dist2 > 4) { set_error("deflate: the match distance must be in [0, 4] (0 = 1)"); return MSL_ERR_ARG; }
    const size_t need = deflate_workspace_bytes(n, container, raw);
    if (!ws || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 15)) {
        set_error("deflate: workspace of %zu bytes (16-byte aligned) needed, %zu given", need, ws_bytes); return MSL_ERR_WORKSPACE;
    }
    ZArgs a;
    memset(&a, 0, sizeof(a));
    a.src = src; a.src_pitch = src_pitch; a.total = total; a.chunk = (unsigned)chunk; a.rows = rows; a.row_bytes = row_bytes;
    a.spv = rows > 0 ? 1u : (unsigned)(total == 0 ? 1 : (total + chunk - 1) / chunk);
    a.prefix = prefix; a.prefix_pitch = prefix_pitch; a.prefix_len = (unsigned)prefix_len; a.expand = expand;
    a.img_w = img_w; a.img_ch = img_ch; a.container = container; a.dist2 = dist2;
    a.meta = reinterpret_cast<uint32_t*>(ws);
    a.slots = reinterpret_cast<uint8_t*>(ws) + (((size_t)n * 16 + 255) & ~(size_t)255);
    a.slot_pitch = deflate_slot_bytes(container, raw);
    {
        // x^(8 * kSeg * k), k = 0 .. 255, and the multipliers of the partial tiles whose length is known here
        struct Tab { uint32_t P[kZThreads]; uint32_t tile[4]; };
        static const Tab tab = [] {
            Tab t;
            const uint32_t step = gf_xpow8(kSeg);
            uint32_t v = 1u << 31;
            for (int k = 0; k < kZThreads; ++k) { t.P[k] = v; v = gf_mul(v, step); }
            for (int k = 0; k < 4; ++k) t.tile[k] = gf_xpow8((unsigned long long)kTile * (k + 1));
            return t;
        }();
        memcpy(a.crcP, tab.P, sizeof(tab.P));
        memcpy(a.crc_tilek, tab.tile, sizeof(tab.tile));
        if (rows == 0 && container != MSL_Z_PNG && chunk % kTile == 0 && chunk <= 4u * kTile && total >= prefix_len + 2 * chunk)
            zero_chunk_stream(container, dist2 ? dist2 : 1, (unsigned)chunk, a.tmpl_meta, reinterpret_cast<uint8_t*>(a.tmpl));
    }
    {
        ProfScope prof(K_DEFLATE, stream);
        deflate_kernel<<<n, kZThreads, 0, stream>>>(a);
        MSL_LAUNCH_CHECK("deflate_kernel");
    }
    {
        ProfScope prof(K_DEFLATE_PACK, stream);
        PackArgs pa;
        pa.slots = a.slots; pa.slot_pitch = a.slot_pitch; pa.meta = a.meta; pa.off = out_off; pa.out = out; pa.out_cap = (unsigned long long)out_cap;
        {
            struct Tab { uint32_t P[kZThreads]; uint32_t x16k; };
            static const Tab tab = [] {
                Tab t;
                const uint32_t step = gf_xpow8(64);
                uint32_t v = 1u << 31;
                for (int k = 0; k < kZThreads; ++k) { t.P[k] = v; v = gf_mul(v, step); }
                t.x16k = gf_xpow8(16384);
                return t;
            }();
            memcpy(pa.P64, tab.P, sizeof(tab.P));
            pa.x16k = tab.x16k;
        }
        pa.n = n;
        pack_kernel<<<n < 148 * 8 ? n : 148 * 8, kZThreads, 0, stream>>>(pa);
        MSL_LAUNCH_CHECK("pack_kernel");
    }
    if (out_meta) MSL_CUDA_CHECK(cudaMemcpyAsync(out_meta, a.meta, (size_t)n * 16, cudaMemcpyDeviceToDevice, stream));
    return MSL_OK;
}

}  // namespace msl
