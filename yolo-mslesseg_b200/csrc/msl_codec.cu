// Byte-stream codec on the device (encode side): RFC 1951 deflate streams inside zlib / gzip / PNG
// containers, so that what crosses PCIe (and what lands on disk) are the compressed files themselves.
//
// Replaces the zlib work behind plt.imsave (scripts/extraer_dataset.py:192,197 -> Pillow PNG encoder), cv2.imwrite
// (utils/utils.py:393, scripts/generar_predicciones.py:153) and nib.save (utils/utils.py:176-177 -> gzip).
//
// deflate_kernel: one CTA per stream (a PNG image, or a 16 KB chunk of a NIfTI file).  Data on this path is either
//   incompressible for a fixed Huffman code (enhanced brain tissue: 8 - 9 bits per literal byte) or one long run
//   (skull-stripped background, masks, padding), so the stream is built from two kinds of blocks decided per WARP:
//   a tile of 16 KB is cut into 256 segments of 64 bytes, one per thread; a byte-SIMD compare gives each thread the mask
//   x[i] == x[i - d] (d = 1: repeated bytes, 4: repeated RGBA pixels / float32 voxels); a ballot tells the warp which of its
//   32 segments are entirely inside a run.  Every maximal group of neighbouring run segments becomes ONE fixed-Huffman
//   block holding a few 258-byte matches (13 bits each); every group of other segments becomes ONE stored block (raw
//   bytes, byte aligned).  A warp's output is therefore whole bytes (a trailing fixed block is realigned by an empty stored
//   block), sizes are known from the ballot alone, a block scan places the warps, the headers / matches are written by
//   lane 0 and the stored bytes are copied by the warp.  No literal is ever Huffman-coded - neither here nor in the decoder,
//   which meets ~10 symbols per 2 KB instead of 2,000.  Staging area, 16-byte flushes, carry to the next tile, Adler-32 /
//   CRC-32 of the raw bytes and the container header / trailer are handled by the same CTA.
// scan_sizes_kernel + pack_kernel: the variable-length streams are packed back to back (exclusive scan of the sizes) into
//   one buffer = one D2H copy; PNG's IDAT CRC-32 (over the compressed bytes) is computed during the copy.
#include <cstring>

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
    uint32_t crc_tile;         // x^(8 * kTile)
    uint32_t crc_part_len[2], crc_part_mul[2];   // x^(8 * t) for the partial-tile lengths t known on the host (0 = unused)
};

__device__ __forceinline__ unsigned hdr_len_of(int container) {
    return (0x2b180200u >> (8 * (container & 3))) & 0xffu;       // raw 0, zlib 2, gzip 24, PNG 43 (no jump table)
}

// ---- bit writer into the zeroed staging area (LSB-first, RFC 1951 section 3.1.1)
struct BitWriter {
    uint32_t* y;          // staging words
    unsigned long long acc;
    int nacc;             // valid bits in acc (the first nacc0 of them are zeros standing for another thread's bits)
    unsigned word;
    bool first;
    __device__ __forceinline__ void init(uint32_t* y_, unsigned bitpos) { y = y_; acc = 0; nacc = (int)(bitpos & 31u); word = bitpos >> 5; first = true; }
    __device__ __forceinline__ void put(uint32_t v, int n) {
        acc |= (unsigned long long)v << nacc;
        nacc += n;
        if (nacc >= 32) {
            if (first) { atomicOr(&y[word], (uint32_t)acc); first = false; }
            else y[word] = (uint32_t)acc;
            acc >>= 32; nacc -= 32; ++word;
        }
    }
    __device__ __forceinline__ void finish() { if (nacc > 0 && (uint32_t)acc) atomicOr(&y[word], (uint32_t)acc); }
};

__device__ __forceinline__ uint32_t rev_bits(uint32_t v, int n) { return __brev(v) >> (32 - n); }

// fixed Huffman code of a literal / length symbol (RFC 1951 section 3.2.6), already bit-reversed; returns the bit count
__device__ __forceinline__ int fixed_litlen(uint32_t sym, uint32_t& code) {
    if (sym < 144) { code = rev_bits(0x30 + sym, 8); return 8; }
    if (sym < 256) { code = rev_bits(0x190 + (sym - 144), 9); return 9; }
    if (sym < 280) { code = rev_bits(sym - 256, 7); return 7; }
    code = rev_bits(0xc0 + (sym - 280), 8); return 8;
}

// cost in bits / emission of one match (length 3..258, distance 1..32768)
template <bool EMIT>
__device__ __forceinline__ int put_match(BitWriter& bw, int len, int dist) {
    uint32_t sym, eb = 0, ev = 0;
    if (len == 258) sym = 285;
    else {
        const uint32_t l = (uint32_t)len - 3;
        if (l < 8) sym = 257 + l;
        else { eb = 29 - __clz(l); sym = 257 + 4 * (eb + 1) + ((l >> eb) & 3); ev = l & ((1u << eb) - 1); }
    }
    const uint32_t D = (uint32_t)dist - 1;
    uint32_t dcode, deb = 0, dev = 0;
    if (D < 4) dcode = D;
    else { deb = 30 - __clz(D); dcode = 2 * (deb + 1) + ((D >> deb) & 1); dev = D & ((1u << deb) - 1); }
    uint32_t code;
    const int n = fixed_litlen(sym, code);
    if (EMIT) {
        bw.put(code, n);
        if (eb) bw.put(ev, (int)eb);
        bw.put(rev_bits(dcode, 5), 5);
        if (deb) bw.put(dev, (int)deb);
    }
    return n + (int)eb + 5 + (int)deb;
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

// `bytes` (a multiple of 64, at most 2048) of one run as matches of up to 258 bytes: returns the bits, emits when EMIT
template <bool EMIT>
__device__ __forceinline__ int put_run(BitWriter& bw, int bytes, int d) {
    int bits = 0;
    while (bytes > 0) {
        int take = bytes < 258 ? bytes : 258;
        if (bytes - take > 0 && bytes - take < 3) take -= 3;       // (cannot happen for multiples of 64; kept for safety)
        bits += put_match<EMIT>(bw, take, d);
        bytes -= take;
    }
    return bits;
}

// ---- one warp's 32 segments as blocks, every lane working for itself
// Blocks alternate between the two kinds, and a stored block ends on a byte boundary, so every fixed block STARTS on one (at
// the warp's first byte or behind a stored block).  That makes every size local: the lane that starts a fixed block emits the
// block and the header of whatever stored block follows it (an empty one at the end of the warp) - whole bytes; the lane that
// starts a stored block emits a header only if it opens the warp, and accounts for the block's bytes.  A warp scan of those
// sizes gives every lane its place; headers are written by their lanes in parallel, and every literal lane copies its own
// 16 bytes.
struct LanePlan {
    bool start, run;
    int bytes;        // bytes of the block this lane starts
    int next_bytes;   // bytes of the stored block that follows this lane's fixed block (0: none, an empty one is emitted)
    int hbits;        // bits of the fixed block (header + matches + end of block)
    int size;         // bytes this lane contributes to the warp's output
};

__device__ __forceinline__ LanePlan lane_plan(unsigned run_mask, int nact, int wbytes, int d, int lane) {
    LanePlan p;
    p.start = false; p.run = false; p.bytes = 0; p.next_bytes = 0; p.hbits = 0; p.size = 0;
    if (lane >= nact) return p;
    const unsigned act = nact == 32 ? FULL : ((1u << nact) - 1u);
    const unsigned runm = run_mask & act;
    const unsigned starts = ((runm ^ (runm << 1)) | 1u) & act;
    p.run = (runm >> lane) & 1u;
    p.start = (starts >> lane) & 1u;
    if (!p.start) return p;
    const unsigned after = lane == 31 ? 0u : (starts >> (lane + 1));
    const int nl = after ? __ffs((int)after) : nact - lane;
    p.bytes = min(kSeg * nl, wbytes - kSeg * lane);
    BitWriter none;
    if (p.run) {
        const int m = lane + nl;
        if (m < nact) {
            const unsigned after2 = m == 31 ? 0u : (starts >> (m + 1));
            const int nl2 = after2 ? __ffs((int)after2) : nact - m;
            p.next_bytes = min(kSeg * nl2, wbytes - kSeg * m);
        }
        p.hbits = 3 + put_run<false>(none, p.bytes, d) + 7;
        p.size = ((p.hbits + 3 + 7) >> 3) + 4;
    } else {
        p.size = (lane == 0 ? 5 : 0) + p.bytes;
    }
    return p;
}

// emission: ybase = the warp's first byte in the staging area, off = this lane's exclusive prefix of the sizes
__device__ __forceinline__ void lane_emit(const LanePlan& p, const SegMasks& m, unsigned run_mask, int nact, int d, uint8_t* Ys,
                                          unsigned ybase, int off, int lane) {
    uint32_t* Y32 = reinterpret_cast<uint32_t*>(Ys);
    if (p.start) {
        BitWriter bw;
        bw.init(Y32, 8u * (ybase + (unsigned)off));
        if (p.run) {
            bw.put(2u, 3);                                         // BFINAL 0, BTYPE 01
            put_run<true>(bw, p.bytes, d);
            bw.put(0u, 7);                                         // end of block
            bw.put(0u, 3 + ((-(p.hbits + 3)) & 7));                // stored header, padding to the byte boundary
            bw.put((uint32_t)p.next_bytes | ((~(uint32_t)p.next_bytes & 0xffffu) << 16), 32);
        } else if (lane == 0) {
            bw.put(0u, 8);
            bw.put((uint32_t)p.bytes | ((~(uint32_t)p.bytes & 0xffffu) << 16), 32);
        }
        bw.finish();
    }
    // literal lanes: the start lane of my block, its offset, my 16 bytes behind it
    const unsigned act = nact == 32 ? FULL : ((1u << nact) - 1u);
    const unsigned runm = run_mask & act;
    const unsigned starts = ((runm ^ (runm << 1)) | 1u) & act;
    const unsigned below = starts & (lane == 31 ? FULL : ((2u << lane) - 1u));
    const int ms = below ? 31 - __clz((int)below) : 0;
    const int off_ms = __shfl_sync(FULL, off, ms);
    if (lane < nact && !p.run && m.len > 0) {
        const unsigned a = ybase + (unsigned)off_ms + (ms == 0 ? 5u : 0u) + (unsigned)(kSeg * (lane - ms));
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

__global__ void __launch_bounds__(kZThreads, 4) deflate_kernel(const ZArgs a) {
    __shared__ __align__(16) uint8_t Xs[kLook + kTile + 4 * (kZThreads + 2) + 16];
    __shared__ __align__(16) uint8_t Ys[kYBytes + 16];
    __shared__ uint32_t crc_table[256];
    __shared__ uint32_t crc_part[kZThreads];
    __shared__ int scan_w[kZThreads / 32];
    __shared__ unsigned long long red_a[kZThreads / 32], red_b[kZThreads / 32];
    __shared__ unsigned s_state[4];     // [0] bits carried in Ys, [1] bytes written to the slot, [2] running CRC
    uint8_t* X = Xs + kLook;
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
    if (want_crc) {
        uint32_t c = (uint32_t)tid;
#pragma unroll
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
        crc_table[tid] = c;
    }
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
        s_state[0] = 8 * hdr; s_state[1] = 0; s_state[2] = 0;          // [0]: bits carried in the staging area (whole bytes)
    }
    block_sync();

    unsigned long long sa = 0, sb = 0;                  // Adler-32 partial sums
    for (unsigned tb = 0; tb < n || tb == 0; tb += kTile) {
        const unsigned tn = min((unsigned)kTile, n - tb);
        // ---- load the tile (image mode inserts the filter byte 0 in front of every scanline)
        if (!image) {
            const unsigned long long g0 = r0 + tb;            // raw offset of the tile inside the source
            if (!a.expand && g0 >= plen && (reinterpret_cast<uintptr_t>(src + (g0 - plen)) & 15) == 0) {
                const uint8_t* body = src + (g0 - plen);
                const uint4* s4 = reinterpret_cast<const uint4*>(body);
                for (unsigned q0_ = 0; q0_ < ((tn >> 4)); q0_ += kZThreads) if (const unsigned q = q0_ + (unsigned)tid; q < ((tn >> 4))) {
                    const uint4 v = __ldg(s4 + q);
                    uint32_t* x32 = reinterpret_cast<uint32_t*>(X + xphys((int)(16 * q)));      // 4-byte aligned (16-byte groups stay inside a segment)
                    x32[0] = v.x; x32[1] = v.y; x32[2] = v.z; x32[3] = v.w;
                }
                if (const unsigned i = (tn & ~15u) + tid; i < tn) XP((int)i) = __ldg(body + i);       // (< 16 bytes)
            } else if (a.expand && g0 >= plen && ((g0 - plen) & 3) == 0) {
                // uint8 mask stored as float32: one mask byte per output word (0.0f / 1.0f)
                const uint8_t* body = src + ((g0 - plen) >> 2);
                for (unsigned q0_ = 0; q0_ < ((tn + 3) >> 2); q0_ += kZThreads) if (const unsigned q = q0_ + (unsigned)tid; q < ((tn + 3) >> 2))
                    *reinterpret_cast<uint32_t*>(X + xphys((int)(4 * q))) = __ldg(body + q) ? 0x3f800000u : 0u;
            } else {
                for (unsigned i0_ = 0; i0_ < (tn); i0_ += kZThreads) if (const unsigned i = i0_ + (unsigned)tid; i < (tn)) {
                    const unsigned long long g = g0 + i;
                    uint8_t v;
                    if (g < plen) v = __ldg(pfx + g);
                    else if (!a.expand) v = __ldg(src + (g - plen));
                    else {
                        const unsigned long long b = g - plen;
                        const uint8_t m = __ldg(src + (b >> 2));
                        v = m ? (uint8_t)(0x3f800000u >> (8 * (unsigned)(b & 3))) : (uint8_t)0;
                    }
                    XP((int)i) = v;
                }
            }
        } else {
            // all of the thread's loads are issued before the first one is used
            uint8_t v[kTile / kZThreads];
#pragma unroll
            for (int j = 0; j < kTile / kZThreads; ++j) {
                const unsigned i = tid + j * kZThreads;
                v[j] = 0;
                if (i < tn) {
                    const unsigned g = tb + i, r = __umulhi(g, rl_magic), c = g - r * rl;
                    if (c) v[j] = __ldg(src + (size_t)r * (rl - 1) + (c - 1));
                }
            }
#pragma unroll
            for (int j = 0; j < kTile / kZThreads; ++j) {
                const unsigned i = tid + j * kZThreads;
                if (i < tn) XP((int)i) = v[j];
            }
        }
        block_sync();
        // ---- the thread's segment as masks; checksums of the raw bytes
        const int beg = tid * kSeg, end = min(beg + kSeg, (int)tn);
        const SegMasks m = seg_masks(X, beg, end, (long long)tb, dist);
        if (want_adler) { sa += m.s1; sb += (unsigned long long)(n - (tb + beg)) * m.s1 - m.s2; }
        if (want_crc && tn > 0) {
            if (__syncwarp(), !__syncthreads_or(m.any != 0 || tb < 4u)) {
                // an all-zero tile: a zero-initialised CRC register stays zero, only the running value moves on
                if (tid == 0) s_state[2] = gf_mul(s_state[2], tn == kTile ? a.crc_tile : (tn == a.crc_part_len[0] ? a.crc_part_mul[0] : (tn == a.crc_part_len[1] ? a.crc_part_mul[1] : gf_xpow8(tn))));
            } else {
                // right-aligned segments: thread t takes the 16 bytes that end (255 - t) segments before the end of the tile, so
                // that short tiles leave the FRONT threads short (a zero register is unchanged by missing leading bytes).  The
                // all-ones initial register of CRC-32 is folded in by complementing the first four bytes of the stream.
                const int e = (int)tn - (kZThreads - 1 - tid) * kSeg, b = e - kSeg;
                uint32_t c = 0;
                for (int i = max(b, 0); i < e; ++i) {
                    uint32_t v = XP(i);
                    if (tb + (unsigned)i < 4u) v ^= 0xffu;
                    c = crc_table[(c ^ v) & 0xffu] ^ (c >> 8);
                }
                // every partial is moved to the end of the tile by ONE multiplication, then the 256 values are XOR-ed
                c = e > 0 ? gf_mul(a.crcP[kZThreads - 1 - tid], c) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(FULL, c, o);
                if (lane == 0) crc_part[warp] = c;
                block_sync();
                if (tid == 0) {
                    uint32_t t = 0;
                    for (int w = 0; w < kZThreads / 32; ++w) t ^= crc_part[w];
                    const uint32_t mul = tn == kTile ? a.crc_tile : (tn == a.crc_part_len[0] ? a.crc_part_mul[0] : (tn == a.crc_part_len[1] ? a.crc_part_mul[1] : gf_xpow8(tn)));
                    s_state[2] = gf_mul(s_state[2], mul) ^ t;
                }
            }
        }
        // ---- blocks per warp: sizes from the ballot, a block scan over the warps, emission
        const bool isrun = m.len == kSeg && m.E == ((1ull << kSeg) - 1ull);
        const unsigned run_mask = __ballot_sync(FULL, isrun);
        const int wbeg = warp * 32 * kSeg;
        const int wbytes = max(0, min(32 * kSeg, (int)tn - wbeg));
        const int nact = (wbytes + kSeg - 1) / kSeg;
        const LanePlan plan = lane_plan(run_mask, nact, wbytes, dist, lane);
        const int lincl = warp_incl_scan(plan.size, lane);
        if (lane == 31) scan_w[warp] = lincl;
        block_sync();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kZThreads / 32; ++w) { const int v = scan_w[w]; if (w < warp) wbase += v; total += v; }
        const unsigned bit0 = s_state[0];
        lane_emit(plan, m, run_mask, nact, dist, Ys, (bit0 >> 3) + (unsigned)wbase, lincl - plan.size, lane);
        total *= 8;
        block_sync();
        unsigned nbits = bit0 + (unsigned)total;
        const bool last = tb + kTile >= n;
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
                    if (n >= 4) chk = s_state[2] ^ 0xffffffffu;
                    else {                                  // streams shorter than the register: the plain definition (the tile is still in X)
                        uint32_t c = 0xffffffffu;
                        for (unsigned i = 0; i < n; ++i) c = crc_table[(c ^ XP((int)i)) & 0xffu] ^ (c >> 8);
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
                const unsigned fsize = s_state[1] + nb;
                uint32_t* m = a.meta + 4 * (size_t)s;
                m[0] = fsize; m[1] = n; m[2] = chk; m[3] = idat_type;
                if (a.container == MSL_Z_GZIP) {
                    // compressed member size and raw size into the 'MS' extra subfield (bytes 16..23 of the member): still in
                    // the staging area for a one-tile stream, already in the slot (flushed before earlier barriers) otherwise
                    uint8_t* h = s_state[1] == 0 ? Ys + 16 : slot + 16;
                    h[0] = fsize; h[1] = fsize >> 8; h[2] = fsize >> 16; h[3] = fsize >> 24;
                    h[4] = n; h[5] = n >> 8; h[6] = n >> 16; h[7] = n >> 24;
                }
                s_state[3] = nb;
            }
            block_sync();
            const unsigned nb = s_state[3], done = s_state[1];
            uint4* d4 = reinterpret_cast<uint4*>(slot + done);
            const uint4* y4 = reinterpret_cast<const uint4*>(Ys);
            for (unsigned q0_ = 0; q0_ < (((nb + 15) >> 4)); q0_ += kZThreads) if (const unsigned q = q0_ + (unsigned)tid; q < (((nb + 15) >> 4))) d4[q] = y4[q];
            break;
        }
        // ---- flush whole 16-byte groups, carry the rest to the front of the staging area
        const unsigned nfl = (nbits >> 3) & ~15u, done = s_state[1];
        {
            uint4* d4 = reinterpret_cast<uint4*>(slot + done);
            const uint4* y4 = reinterpret_cast<const uint4*>(Ys);
            for (unsigned q0_ = 0; q0_ < ((nfl >> 4)); q0_ += kZThreads) if (const unsigned q = q0_ + (unsigned)tid; q < ((nfl >> 4))) d4[q] = y4[q];
        }
        uint4 carry = make_uint4(0, 0, 0, 0);
        if (tid == 0) carry = *reinterpret_cast<const uint4*>(Ys + nfl);          // < 16 bytes and a partial byte remain
        if (tid == 1) carry = *reinterpret_cast<const uint4*>(Ys + nfl + 16);
        block_sync();
        for (int q0_ = 0; q0_ < ((kYBytes + 16) / 4); q0_ += kZThreads) if (const int q = q0_ + (int)tid; q < ((kYBytes + 16) / 4)) Y32[q] = 0;
        // the look-back of the next tile = the tail of this one
        uint32_t lb = 0;
        if (tid < kLook / 4) lb = *reinterpret_cast<const uint32_t*>(X + xphys(kTile - kLook + 4 * tid));
        block_sync();
        if (tid == 0) { *reinterpret_cast<uint4*>(Ys) = carry; s_state[0] = nbits - 8 * nfl; s_state[1] = done + nfl; }
        if (tid == 1) *reinterpret_cast<uint4*>(Ys + 16) = carry;
        if (tid < kLook / 4) reinterpret_cast<uint32_t*>(Xs)[tid] = lb;
        block_sync();
    }
}

// exclusive scan of the stream sizes -> byte offsets of the packed streams (one CTA; n is a few ten thousand at most)
__global__ void __launch_bounds__(1024) scan_sizes_kernel(const uint32_t* meta, int n, unsigned align, unsigned long long* off) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    block_sync();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        unsigned long long v = i < n ? (((unsigned long long)meta[4 * (size_t)i] + align - 1) / align) * align : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        block_sync();
        unsigned long long wb = 0, tot = 0;
        for (int w = 0; w < 32; ++w) { if (w < warp) wb += wsum[w]; tot += wsum[w]; }
        const unsigned long long c = carry_s;
        if (i < n) off[i] = c + wb + inc - v;
        block_sync();
        if (tid == 0) carry_s = c + tot;
        block_sync();
    }
    if (tid == 0) off[n] = carry_s;
}

// copies stream s from its slot to out + off[s]; PNG: IDAT length and CRC-32 over (type + compressed data)
struct PackArgs {
    const uint8_t* slots;
    size_t slot_pitch;
    const uint32_t* meta;
    const unsigned long long* off;
    uint8_t* out;
    unsigned long long out_cap;
    uint32_t P64[kZThreads];   // x^(8 * 64 * k)
    uint32_t x16k;             // x^(8 * 16384)
};

__global__ void __launch_bounds__(kZThreads) pack_kernel(const PackArgs a) {
    __shared__ uint32_t crc_table[256];
    __shared__ uint32_t part[kZThreads / 32];
    __shared__ uint32_t s_run;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned s = blockIdx.x;
    const uint8_t* slot = a.slots + (size_t)s * a.slot_pitch;
    const uint32_t fsize = a.meta[4 * (size_t)s], idat = a.meta[4 * (size_t)s + 3];
    const unsigned long long o = a.off[s];
    if (o + fsize > a.out_cap) return;                      // the caller compares off[n] with the capacity
    uint8_t* dst = a.out + o;
    uint32_t clen = 0;                                      // PNG: bytes the IDAT CRC covers = "IDAT" + zlib stream
    if (idat) {
        clen = fsize - idat - 16;                           // file = ... [idat - 4: length][idat: "IDAT" + data][CRC][IEND chunk: 12]
        uint32_t c = (uint32_t)tid;
#pragma unroll
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
        crc_table[tid] = c;
        if (tid == 0) s_run = 0;
    }
    // aligned body: destination words, source read byte-wise when the packed offset is not aligned
    const unsigned head = (unsigned)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
    const unsigned h = min(head, fsize);
    if (tid < (int)h) dst[tid] = slot[tid];
    const unsigned nwords = (fsize - h) >> 2;
    if (((reinterpret_cast<uintptr_t>(slot) + h) & 3) == 0) {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(slot + h);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + h);
        for (unsigned q0_ = 0; q0_ < (nwords); q0_ += kZThreads) if (const unsigned q = q0_ + (unsigned)tid; q < (nwords)) d32[q] = s32[q];
    } else {
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + h);
        for (unsigned q0_ = 0; q0_ < (nwords); q0_ += kZThreads) if (const unsigned q = q0_ + (unsigned)tid; q < (nwords)) {
            const uint8_t* p = slot + h + 4 * q;
            d32[q] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        }
    }
    if (const unsigned i = h + 4 * nwords + tid; i < fsize) dst[i] = slot[i];                       // (< 4 bytes)
    if (!idat) return;
    block_sync();
    // CRC-32 of the IDAT chunk in tiles of 16 KB that are aligned to the END of the chunk (only the first tile is short, and a
    // zero register does not see missing leading bytes): a thread takes 64 bytes, moves its value to the end of the tile with
    // one multiplication, the tile's XOR is folded into the running value.  The all-ones initial register = the first four
    // bytes complemented.
    const uint8_t* C = slot + idat;
    const int ntile = (int)((clen + 16383u) >> 14);
    for (int j = 0; j < ntile; ++j) {
        const long long te = (long long)clen - (long long)(ntile - 1 - j) * 16384;
        const long long e = te - (long long)(kZThreads - 1 - tid) * 64, b = e - 64;
        uint32_t c = 0;
        const long long lo = max(b, max(te - 16384, 0ll));
        for (long long i = lo; i < e; ++i) {
            uint32_t v = C[i];
            if (i < 4) v ^= 0xffu;
            c = crc_table[(c ^ v) & 0xffu] ^ (c >> 8);
        }
        c = e > lo ? gf_mul(a.P64[kZThreads - 1 - tid], c) : 0u;
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

inline size_t raw_len_of(size_t chunk, int rows, int row_bytes) { return rows > 0 ? (size_t)rows * ((size_t)row_bytes + 1) : chunk; }

}  // namespace

size_t deflate_slot_bytes(int container, size_t raw) {
    const size_t hdr = container == MSL_Z_ZLIB ? 2 : container == MSL_Z_GZIP ? 24 : container == MSL_Z_PNG ? 43 : 0;
    return ((hdr + (raw * 9 + 10 + 7) / 8 + 32 + 16) + 15) & ~(size_t)15;
}

size_t deflate_workspace_bytes(int n, int container, size_t raw) {
    return (size_t)n * deflate_slot_bytes(container, raw) + (((size_t)n * 16 + 255) & ~(size_t)255);
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
        struct Tab { uint32_t P[kZThreads]; uint32_t tile; };
        static const Tab tab = [] {
            Tab t;
            const uint32_t step = gf_xpow8(kSeg);
            uint32_t v = 1u << 31;
            for (int k = 0; k < kZThreads; ++k) { t.P[k] = v; v = gf_mul(v, step); }
            t.tile = gf_xpow8(kTile);
            return t;
        }();
        memcpy(a.crcP, tab.P, sizeof(tab.P));
        a.crc_tile = tab.tile;
        const size_t lens[2] = {raw % kTile, rows > 0 ? 0 : (total % (chunk ? chunk : 1)) % kTile};
        for (int k = 0; k < 2; ++k) { a.crc_part_len[k] = (uint32_t)lens[k]; a.crc_part_mul[k] = lens[k] ? gf_xpow8(lens[k]) : 0u; }
    }
    {
        ProfScope prof(K_DEFLATE, stream);
        deflate_kernel<<<n, kZThreads, 0, stream>>>(a);
        MSL_LAUNCH_CHECK("deflate_kernel");
    }
    {
        ProfScope prof(K_DEFLATE_SCAN, stream);
        scan_sizes_kernel<<<1, 1024, 0, stream>>>(a.meta, n, 1u, out_off);
        MSL_LAUNCH_CHECK("scan_sizes_kernel");
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
        pack_kernel<<<n, kZThreads, 0, stream>>>(pa);
        MSL_LAUNCH_CHECK("pack_kernel");
    }
    if (out_meta) MSL_CUDA_CHECK(cudaMemcpyAsync(out_meta, a.meta, (size_t)n * 16, cudaMemcpyDeviceToDevice, stream));
    return MSL_OK;
}

}  // namespace msl
