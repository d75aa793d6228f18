// Byte-stream codec on the device (encode side): RFC 1951 deflate streams inside zlib / gzip / PNG
// containers, so that what crosses PCIe (and what lands on disk) are the compressed files themselves.
//
// Replaces the zlib work behind plt.imsave (scripts/extraer_dataset.py:192,197 -> Pillow PNG encoder), cv2.imwrite
// (utils/utils.py:393, scripts/generar_predicciones.py:153) and nib.save (utils/utils.py:176-177 -> gzip).
//
// deflate_kernel: one CTA per stream (a PNG image, or a 64 KB chunk of a NIfTI file).  The stream is ONE fixed-Huffman
//   block.  A tile of 16 KB is tokenised by 256 threads, 64 bytes each: greedy choice between a literal and a run match at
//   distance 1 (runs: skull-stripped background, masks) or at a second distance (4: RGBA pixels / float32 voxels); matches
//   never cross a thread's segment, except that four neighbouring segments that are entirely one distance-1 run are merged
//   into a single 256-byte match (background costs 13 bits per 256 bytes).  Pass A counts bits, a block scan turns them
//   into bit offsets, pass B re-tokenises and ORs the codes into a zeroed shared-memory staging area that is streamed out
//   16 bytes at a time; partial bytes carry over to the next tile.  Adler-32 / CRC-32 of the raw bytes are accumulated from
//   the same shared-memory tile.  The container header and trailer are written by the same CTA.
// scan_sizes_kernel + pack_kernel: the variable-length streams are packed back to back (exclusive scan of the sizes) into
//   one buffer = one D2H copy; PNG's IDAT CRC-32 (over the compressed bytes) is computed during the copy.
#include <cstring>

#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

namespace {

constexpr int kZThreads = 256;
constexpr int kSeg = 64;                         // bytes per thread per tile
constexpr int kTile = kZThreads * kSeg;          // 16 KB
constexpr int kLook = 16;                        // look-back kept in front of the tile (>= the largest match distance)
constexpr int kYBytes = 16 + kTile * 9 / 8 + 80; // staging: carry + worst case (9 bits / byte) + EOB, trailer, padding
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
    int dist2;                 // second match distance (0 = none)
    uint8_t* slots;            // [n][slot_pitch]
    size_t slot_pitch;
    uint32_t* meta;            // [n][4]: container bytes in the slot, raw bytes, checksum of the raw bytes, offset of the IDAT chunk type
    uint32_t crcM[8];          // x^(8 * 64 * 2^l): combines the per-thread CRCs of a tile
    uint32_t crc_tile;         // x^(8 * kTile)
};

__device__ __forceinline__ unsigned hdr_len_of(int container) {
    return container == MSL_Z_ZLIB ? 2u : container == MSL_Z_GZIP ? 24u : container == MSL_Z_PNG ? 43u : 0u;
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

// greedy tokenisation of X[beg, end): literal or run match at distance 1 / dist2.  gfirst: global raw index of X[0]
// (a match may not reach in front of the stream).  Returns the bit count; *full = the whole 64-byte segment is one
// distance-1 run.
template <bool EMIT>
__device__ __forceinline__ int encode_segment(const uint8_t* X, int beg, int end, long long gfirst, int dist2, BitWriter& bw, bool* full) {
    int bits = 0, i = beg;
    bool isfull = false;
    while (i < end) {
        const int maxl = min(end - i, 258);
        int l1 = 0, l2 = 0;
        const uint32_t x0 = XP(i);
        if (gfirst + i >= 1 && x0 == XP(i - 1)) {
            l1 = 1;
            while (l1 < maxl && XP(i + l1) == x0) ++l1;
        }
        if (dist2 && l1 < maxl && gfirst + i >= dist2 && x0 == XP(i - dist2)) {
            l2 = 1;
            while (l2 < maxl && XP(i + l2) == XP(i + l2 - dist2)) ++l2;
        }
        const int best = l2 > l1 ? l2 : l1, d = l2 > l1 ? dist2 : 1;
        if (best >= 3) {
            if (i == beg && best == end - beg && d == 1 && best == kSeg) isfull = true;
            bits += put_match<EMIT>(bw, best, d);
            i += best;
        } else {
            uint32_t code;
            const int n = fixed_litlen(x0, code);
            if (EMIT) bw.put(code, n);
            bits += n;
            ++i;
        }
    }
    if (full) *full = isfull;
    return bits;
}

__global__ void __launch_bounds__(kZThreads) deflate_kernel(const ZArgs a) {
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

    if (want_crc) {
        uint32_t c = (uint32_t)tid;
#pragma unroll
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
        crc_table[tid] = c;
    }
    for (int q = tid; q < (kYBytes + 16) / 4; q += kZThreads) Y32[q] = 0;
    if (tid < kLook / 4) reinterpret_cast<uint32_t*>(Xs)[tid] = 0;
    __syncthreads();
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
            Ys[24] = 8; Ys[25] = a.img_ch == 4 ? 6 : (a.img_ch == 3 ? 2 : (a.img_ch == 2 ? 4 : 0)); Ys[26] = 0; Ys[27] = 0; Ys[28] = 0;
            uint32_t c = 0xffffffffu;
            for (int i = 12; i < 29; ++i) { c ^= Ys[i]; for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1; }
            c = ~c;
            Ys[29] = c >> 24; Ys[30] = c >> 16; Ys[31] = c >> 8; Ys[32] = c;
            Ys[37] = 'I'; Ys[38] = 'D'; Ys[39] = 'A'; Ys[40] = 'T';       // [33, 37): IDAT length, patched by pack_kernel
            Ys[41] = 0x78; Ys[42] = 0x01;
        }
        Ys[hdr] = 3;                                   // BFINAL = 1, BTYPE = 01 (fixed Huffman), LSB first
        s_state[0] = 8 * hdr + 3; s_state[1] = 0; s_state[2] = 0;
    }
    __syncthreads();

    unsigned long long sa = 0, sb = 0;                  // Adler-32 partial sums
    for (unsigned tb = 0; tb < n || tb == 0; tb += kTile) {
        const unsigned tn = min((unsigned)kTile, n - tb);
        // ---- load the tile (image mode inserts the filter byte 0 in front of every scanline)
        if (!image) {
            const unsigned long long g0 = r0 + tb;            // raw offset of the tile inside the source
            if (!a.expand && g0 >= plen && (reinterpret_cast<uintptr_t>(src + (g0 - plen)) & 15) == 0) {
                const uint8_t* body = src + (g0 - plen);
                const uint4* s4 = reinterpret_cast<const uint4*>(body);
                for (unsigned q = tid; q < (tn >> 4); q += kZThreads) {
                    const uint4 v = __ldg(s4 + q);
                    uint32_t* x32 = reinterpret_cast<uint32_t*>(X + xphys((int)(16 * q)));      // 4-byte aligned (16-byte groups stay inside a segment)
                    x32[0] = v.x; x32[1] = v.y; x32[2] = v.z; x32[3] = v.w;
                }
                for (unsigned i = (tn & ~15u) + tid; i < tn; i += kZThreads) XP((int)i) = __ldg(body + i);
            } else {
                for (unsigned i = tid; i < tn; i += kZThreads) {
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
            for (unsigned i = tid; i < tn; i += kZThreads) {
                const unsigned g = tb + i, r = __umulhi(g, rl_magic), c = g - r * rl;
                XP((int)i) = c == 0 ? (uint8_t)0 : __ldg(src + (size_t)r * (rl - 1) + (c - 1));
            }
        }
        __syncthreads();
        // ---- checksums of the raw bytes
        const int beg = tid * kSeg, end = min(beg + kSeg, (int)tn);
        if (want_adler) {
            for (int i = beg; i < end; ++i) { const unsigned v = XP(i); sa += v; sb += (unsigned long long)(n - (tb + i)) * v; }
        }
        if (want_crc && tn > 0) {
            // right-aligned segments: thread t takes the 64 bytes that end (255 - t) segments before the end of the tile, so
            // that short tiles leave the FRONT threads short (a zero register is unchanged by missing leading bytes)
            const int e = (int)tn - (kZThreads - 1 - tid) * kSeg, b = e - kSeg;
            uint32_t c = 0;
            for (int i = max(b, 0); i < e; ++i) c = crc_table[(c ^ XP(i)) & 0xffu] ^ (c >> 8);
            crc_part[tid] = c;
            __syncthreads();
#pragma unroll 1
            for (int l = 0; l < 8; ++l) {
                const int st = 1 << l;
                if (tid < (kZThreads >> (l + 1))) {
                    const int left = (2 * tid + 1) * st - 1, right = (2 * tid + 2) * st - 1;
                    crc_part[right] = gf_mul(a.crcM[l], crc_part[left]) ^ crc_part[right];
                }
                __syncthreads();
            }
            if (tid == 0) s_state[2] = gf_mul(s_state[2], tn == kTile ? a.crc_tile : gf_xpow8(tn)) ^ crc_part[kZThreads - 1];
        }
        // ---- pass A: bits per thread; four neighbouring all-run segments become one 256-byte match
        BitWriter bw;
        bool full = false;
        int bits = beg < end ? encode_segment<false>(X, beg, end, (long long)tb, a.dist2, bw, &full) : 0;
        const unsigned fm = __ballot_sync(FULL, full);
        const bool merged = ((fm >> (lane & ~3)) & 0xfu) == 0xfu;
        if (merged) bits = (lane & 3) == 0 ? put_match<false>(bw, 4 * kSeg, 1) : 0;
        // block-wide exclusive scan of the bit counts
        int incl = warp_incl_scan(bits, lane);
        if (lane == 31) scan_w[warp] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kZThreads / 32; ++w) { const int v = scan_w[w]; if (w < warp) wbase += v; total += v; }
        const unsigned bit0 = s_state[0];
        // ---- pass B: emit
        if (bits > 0) {
            bw.init(Y32, bit0 + (unsigned)(wbase + incl - bits));
            if (merged) put_match<true>(bw, 4 * kSeg, 1);
            else encode_segment<true>(X, beg, end, (long long)tb, a.dist2, bw, nullptr);
            bw.finish();
        }
        __syncthreads();
        unsigned nbits = bit0 + (unsigned)total;
        const bool last = tb + kTile >= n;
        if (last) {
            // end of block (symbol 256: seven zero bits), byte alignment, trailer
            if (want_adler) {
                for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(FULL, sa, o); sb += __shfl_xor_sync(FULL, sb, o); }
                if (lane == 0) { red_a[warp] = sa; red_b[warp] = sb; }
            }
            __syncthreads();
            if (tid == 0) {
                nbits += 7;
                unsigned nb = (nbits + 7) >> 3;
                uint32_t chk = 0;
                if (want_adler) {
                    unsigned long long ta = 1, tbb = n;
                    for (int w = 0; w < kZThreads / 32; ++w) { ta += red_a[w]; tbb += red_b[w]; }
                    chk = (uint32_t)((tbb % 65521ull) << 16) | (uint32_t)(ta % 65521ull);
                    Ys[nb] = chk >> 24; Ys[nb + 1] = chk >> 16; Ys[nb + 2] = chk >> 8; Ys[nb + 3] = chk;
                    nb += 4;
                } else if (want_crc) {
                    chk = s_state[2] ^ gf_mul(gf_xpow8(n), 0xffffffffu) ^ 0xffffffffu;
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
            __syncthreads();
            const unsigned nb = s_state[3], done = s_state[1];
            uint4* d4 = reinterpret_cast<uint4*>(slot + done);
            const uint4* y4 = reinterpret_cast<const uint4*>(Ys);
            for (unsigned q = tid; q < ((nb + 15) >> 4); q += kZThreads) d4[q] = y4[q];
            break;
        }
        // ---- flush whole 16-byte groups, carry the rest to the front of the staging area
        const unsigned nfl = (nbits >> 3) & ~15u, done = s_state[1];
        {
            uint4* d4 = reinterpret_cast<uint4*>(slot + done);
            const uint4* y4 = reinterpret_cast<const uint4*>(Ys);
            for (unsigned q = tid; q < (nfl >> 4); q += kZThreads) d4[q] = y4[q];
        }
        uint4 carry = make_uint4(0, 0, 0, 0);
        if (tid == 0) carry = *reinterpret_cast<const uint4*>(Ys + nfl);          // < 16 bytes and a partial byte remain
        if (tid == 1) carry = *reinterpret_cast<const uint4*>(Ys + nfl + 16);
        __syncthreads();
        for (int q = tid; q < (kYBytes + 16) / 4; q += kZThreads) Y32[q] = 0;
        // the look-back of the next tile = the tail of this one
        uint32_t lb = 0;
        if (tid < kLook / 4) lb = *reinterpret_cast<const uint32_t*>(X + xphys(kTile - kLook + 4 * tid));
        __syncthreads();
        if (tid == 0) { *reinterpret_cast<uint4*>(Ys) = carry; s_state[0] = nbits - 8 * nfl; s_state[1] = done + nfl; }
        if (tid == 1) *reinterpret_cast<uint4*>(Ys + 16) = carry;
        if (tid < kLook / 4) reinterpret_cast<uint32_t*>(Xs)[tid] = lb;
        __syncthreads();
    }
}

// exclusive scan of the stream sizes -> byte offsets of the packed streams (one CTA; n is a few ten thousand at most)
__global__ void __launch_bounds__(1024) scan_sizes_kernel(const uint32_t* meta, int n, unsigned align, unsigned long long* off) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        unsigned long long v = i < n ? (((unsigned long long)meta[4 * (size_t)i] + align - 1) / align) * align : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        unsigned long long wb = 0, tot = 0;
        for (int w = 0; w < 32; ++w) { if (w < warp) wb += wsum[w]; tot += wsum[w]; }
        const unsigned long long c = carry_s;
        if (i < n) off[i] = c + wb + inc - v;
        __syncthreads();
        if (tid == 0) carry_s = c + tot;
        __syncthreads();
    }
    if (tid == 0) off[n] = carry_s;
}

// copies stream s from its slot to out + off[s]; PNG: IDAT length and CRC-32 over (type + compressed data)
__global__ void __launch_bounds__(kZThreads) pack_kernel(const uint8_t* slots, size_t slot_pitch, const uint32_t* meta,
                                                         const unsigned long long* off, uint8_t* out, unsigned long long out_cap) {
    __shared__ uint32_t crc_table[256];
    __shared__ uint32_t part[kZThreads];
    __shared__ uint32_t M[8];
    __shared__ uint32_t s_k;
    const int tid = threadIdx.x;
    const unsigned s = blockIdx.x;
    const uint8_t* slot = slots + (size_t)s * slot_pitch;
    const uint32_t fsize = meta[4 * (size_t)s], idat = meta[4 * (size_t)s + 3];
    const unsigned long long o = off[s];
    if (o + fsize > out_cap) return;                        // the caller compares off[n] with the capacity
    uint8_t* dst = out + o;
    uint32_t clen = 0;                                      // PNG: bytes the IDAT CRC covers = "IDAT" + zlib stream
    if (idat) {
        clen = fsize - idat - 16;                           // file = ... [idat - 4: length][idat: "IDAT" + data][CRC][IEND chunk: 12]
        uint32_t c = (uint32_t)tid;
#pragma unroll
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
        crc_table[tid] = c;
        if (tid == 0) {
            const uint32_t K = (clen + kZThreads - 1) / kZThreads;
            s_k = K;
            uint32_t m = gf_xpow8(K);
            for (int l = 0; l < 8; ++l) { M[l] = m; m = gf_mul(m, m); }
        }
    }
    // aligned body: destination words, source read byte-wise (the packed offsets are not aligned)
    const unsigned head = (unsigned)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
    const unsigned h = min(head, fsize);
    if (tid < (int)h) dst[tid] = slot[tid];
    const unsigned nwords = (fsize - h) >> 2;
    if (((reinterpret_cast<uintptr_t>(slot) + h) & 3) == 0) {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(slot + h);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + h);
        for (unsigned q = tid; q < nwords; q += kZThreads) d32[q] = s32[q];
    } else {
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + h);
        for (unsigned q = tid; q < nwords; q += kZThreads) {
            const uint8_t* p = slot + h + 4 * q;
            d32[q] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        }
    }
    for (unsigned i = h + 4 * nwords + tid; i < fsize; i += kZThreads) dst[i] = slot[i];
    if (!idat) return;
    __syncthreads();
    // CRC-32 of the IDAT chunk: every thread takes K bytes, right-aligned (see deflate_kernel)
    const uint32_t K = s_k;
    const uint8_t* C = slot + idat;
    {
        const long long start = (long long)clen - (long long)(kZThreads - tid) * K;
        uint32_t c = 0;
        for (long long i = start < 0 ? 0 : start; i < start + (long long)K; ++i) c = crc_table[(c ^ C[i]) & 0xffu] ^ (c >> 8);
        part[tid] = c;
    }
    __syncthreads();
#pragma unroll 1
    for (int l = 0; l < 8; ++l) {
        const int st = 1 << l;
        if (tid < (kZThreads >> (l + 1))) {
            const int left = (2 * tid + 1) * st - 1, right = (2 * tid + 2) * st - 1;
            part[right] = gf_mul(M[l], part[left]) ^ part[right];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t crc = part[kZThreads - 1] ^ gf_mul(gf_xpow8(clen), 0xffffffffu) ^ 0xffffffffu;
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
    if (dist2 < 0 || dist2 > kLook) { set_error("deflate: second match distance must be in [0, %d]", kLook); return MSL_ERR_ARG; }
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
    uint32_t m = gf_xpow8(kSeg);
    for (int l = 0; l < 8; ++l) { a.crcM[l] = m; m = gf_mul(m, m); }
    a.crc_tile = gf_xpow8(kTile);
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
        pack_kernel<<<n, kZThreads, 0, stream>>>(a.slots, a.slot_pitch, a.meta, out_off, out, (unsigned long long)out_cap);
        MSL_LAUNCH_CHECK("pack_kernel");
    }
    if (out_meta) MSL_CUDA_CHECK(cudaMemcpyAsync(out_meta, a.meta, (size_t)n * 16, cudaMemcpyDeviceToDevice, stream));
    return MSL_OK;
}

}  // namespace msl
