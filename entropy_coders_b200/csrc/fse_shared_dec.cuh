// fse_shared_dec.cuh -- decode with one decode table OWNED BY THE CTA and replicated across the shared-memory banks.
//
// The private-table decoders pay 3.35 + 3.19 wavefronts per symbol for the entry and the symbol look-up.  Here the
// reference's 32-bit DecodeTransform (fse.rs:260-265) is one load again and the table is stored R times, copy r in the
// banks {r, r + R, ...}: entry c of copy r is the word at  tabR + (c * R + r) * 4  (linear, one IMAD per address).
//   table_log <= 10: R = 32, lane l reads copy l: exactly one wavefront per look-up;
//   table_log == 11: R = 16, lanes l and l + 16 share copy l & 15, which spans banks l & 15 and (l & 15) + 16
//                    (cell parity picks the bank): at most two wavefronts per look-up.
// 128 KiB either way.  The entry is re-packed for the state chain: the state is kept as the shared-memory ADDRESS of
// its row, and an entry holds the row address of its new_state, so a transition is
//   e = lds(a);  nb = e >> 28;  a' = (e & 0xffff) * 4 + lane_part + (window & ~(~0 << nb)) * 4R      (two IMADs)
// (new_state + bits is an OR in the reference arithmetic: new_state is a multiple of 1 << num_bits).
#pragma once
#include "fse_decode128c.cuh"

namespace fsed {

// R copies (32, 16 or 8; R = 8: four lanes per copy, cell & 3 picks one of its four banks, ~2.9 wavefronts per look-up)
__host__ __device__ inline uint32_t sh_dec_shift(uint32_t R) { return R == 32 ? 7u : (R == 16 ? 6u : 5u); }   // log2(4 * R)
__host__ __device__ inline uint32_t sh_dec_copies_for(uint32_t log2) { return log2 <= 10 ? 32u : 16u; }           // 128 KiB of table
constexpr uint32_t SH_DEC_RING_BYTES = 1024 + 16 + 16;      // 256 words + 2 mirror words (+ pad) + one mbarrier
struct ShDecLayout { uint32_t tab, ring, total; };
__host__ __device__ inline ShDecLayout sh_dec_layout(uint32_t log2, uint32_t R, int warps)
{
    ShDecLayout l;
    l.tab = 0;
    l.ring = (1u << log2) * 4u * R;
    l.total = l.ring + SH_DEC_RING_BYTES * (uint32_t)warps + 128;     // + alignment slack
    return l;
}

// entry = WORD address of the row of new_state (16 bits: shared memory is < 256 KiB) | symbol << 16 | num_bits << 28: the
// symbol is a whole byte (one PRMT gathers the four symbols of a lane, no shifts) and num_bits is one shift away
__device__ __forceinline__ uint32_t shd_entry(uint32_t tab_saddr, uint32_t sh, uint32_t new_state, uint32_t sym, uint32_t nb)
{
    return ((tab_saddr + (new_state << sh)) >> 2) | (sym << 16) | (nb << 28);
}
// the CTA copies a decode table in the reference layout (new_state | symbol << 16 | num_bits << 24) into the replicated form
__device__ __forceinline__ void sh_replicate_dec(const uint32_t *__restrict__ tab, uint32_t log2, uint32_t R, uint8_t *tabR, uint32_t tab_saddr,
                                                 int tid, int nthr)
{
    const uint32_t sh = sh_dec_shift(R);
    const uint32_t vsh = sh - 4;                              // log2(16-byte vectors per row)
    const uint32_t nvec = (1u << log2) << vsh;
    for (uint32_t i = tid; i < nvec; i += nthr) {
        const uint32_t r = tab[i >> vsh];
        const uint32_t w = shd_entry(tab_saddr, sh, r & 0xffffu, (r >> 16) & 0xffu, r >> 24);
        reinterpret_cast<uint4 *>(tabR)[i] = make_uint4(w, w, w, w);
    }
}

// the next row address: (word address of new_state's row) * 4 + lane part, plus the bits read * row size: one LOP + two IMADs
template <int SH>
__device__ __forceinline__ uint32_t shd_next(uint32_t e, uint32_t win, uint32_t n, uint32_t lanepart)
{
    uint32_t a;
    asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(a) : "r"(e & 0xffffu), "r"(lanepart));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a) : "r"(win & ~(0xffffffffu << n)), "n"(1 << SH));
    return a;
}

struct ShDecWarp {
    uint32_t ring_saddr, bar, par;
    uint32_t *ring;
};

// One stream of 128 interleaved states (payload only) by one warp against the CTA's table.  Returns the status.
// lanepart = (lane & (R - 1)) * 4, row0 = tab_saddr.
template <int SH>
__device__ __forceinline__ int sh_decode_payload_warp(const uint8_t *pay, uint32_t plen, uint8_t *out, uint32_t bn, uint32_t log2,
                                                      uint32_t row0, uint32_t lanepart, ShDecWarp &wk, int lane)
{
    constexpr uint32_t sh = SH;
    const uint32_t N = 128;
    if (plen == 0 || pay[plen - 1] == 0) return ST_NO_MARKER;
    const uint32_t bias = (uint32_t)((uintptr_t)pay & 15);     // bulk copies need 16-byte aligned global chunks
    const uint32_t *origin = reinterpret_cast<const uint32_t *>(pay - bias);
    uint32_t cur = (plen - 1) * 8 + ilog2u(pay[plen - 1]) + 8 * bias;
    const uint32_t floor_bits = 8 * bias;
    if (cur - floor_bits < N * log2) return ST_LENGTH;
    const uint32_t topq = cur >> 5;
    uint32_t lowq = (topq & ~127u) >= 128 ? (topq & ~127u) - 128 : 0;
    uint32_t *ring = wk.ring;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t w = lowq + lane + 32 * k;
        if (w <= topq) {
            uint32_t x = __ldg(origin + w);
            ring[w & 255] = x;
            if ((w & 255) < 2) ring[256 + (w & 255)] = x;       // mirror: ring[256..257] == ring[0..1]
        }
    }
    __syncwarp();
    const uint32_t ring_saddr = wk.ring_saddr;
    auto ring_bits64 = [&](uint32_t q, uint32_t &lo, uint32_t &hi) {
        uint32_t ad = ring_saddr + ((q >> 3) & 0x3fcu);
        uint32_t w0, w1, w2;
        asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                     : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(ad));
        uint32_t s = q & 31;
        lo = __funnelshift_r(w0, w1, s);
        hi = __funnelshift_r(w1, w2, s);
    };
    uint32_t pending = 0, tma_ok = 1;                        // 32-bit flags: touched on the rare paths only, no bool packing per round
    // The ring holds words [lowq, lowq + 256).  One test per round covers both stages of the refill: nothing to do while
    // the cursor is in the upper half.  Below it, one lane starts a 512-byte bulk copy (TMA) of the next lower 128 words
    // into the dead upper half, and the warp waits on the mbarrier only when a round could reach below lowq.
    auto stage = [&]() {
        if ((cur >> 5) + 3 >= lowq + 128 || !lowq) return;
        if (!pending) {
            __syncwarp();
            if (lane == 0) bulk_g2s(ring_saddr + (((lowq - 128) & 255) << 2), origin + (lowq - 128), 512, wk.bar);
            pending = 1;
        }
        if ((cur >> 5) < lowq + 56) {                          // a round takes at most 52 words
            if (!mbar_wait(wk.bar, wk.par)) tma_ok = 0;
            wk.par ^= 1;
            lowq -= 128;
            pending = 0;
            if ((lowq & 255) == 0) {
                __syncwarp();
                if (lane < 2) ring[256 + lane] = ring[lane];
            }
            __syncwarp();
        }
    };
    const uint32_t lrow = row0 + lanepart;
    // Decoder::new, fse.rs:349-352: states are read 0, 1, 2, ... from the top of the stack
    uint32_t a0, a1, a2, a3;
    {
        const uint32_t m = (1u << log2) - 1u;
        uint32_t lo, hi;
        ring_bits64(cur - (4 * lane + 4) * log2, lo, hi);
        const uint64_t w = ((uint64_t)hi << 32) | lo;
        a3 = lrow + (((uint32_t)w & m) << sh);
        a2 = lrow + (((uint32_t)(w >> log2) & m) << sh);
        a1 = lrow + (((uint32_t)(w >> (2 * log2)) & m) << sh);
        a0 = lrow + (((uint32_t)(w >> (3 * log2)) & m) << sh);
    }
    cur -= N * log2;
    const uint32_t body = bn - N;
    const bool out_aligned = (((uintptr_t)out) & 3) == 0;
    uint32_t bad = 0;
    uint32_t i0 = 0;
    // next row address: the row of new_state (in the entry) | lane part, plus the bits read; the product is an IMAD (FMA pipe)
#define SHD_NEXT(e, win, n) shd_next<SH>(e, win, n, lanepart)
#define SHD_NB(e) ((e) >> 28)                               /* (IMAD.HI runs at a quarter of the ALU rate: tools/pipe_bench.cu) */
#define SHD_SYM(e) (((e) >> 16) & 0xffu)
    // one full round: fse.rs:363-373 on four chains per lane; STORE writes the lane's four symbols
#define SHD_ROUND(STORE)                                                                                              \
    {                                                                                                                 \
        stage();                                                                                                      \
        const uint32_t e0 = lds_u32(a0), e1 = lds_u32(a1), e2 = lds_u32(a2), e3 = lds_u32(a3);                         \
        const uint32_t n0 = SHD_NB(e0), n1 = SHD_NB(e1), n2 = SHD_NB(e2), n3 = SHD_NB(e3);                              \
        const uint32_t n23 = n2 + n3, n123 = n1 + n23, nbs = n0 + n123;                                                 \
        const uint32_t incl = warp_incl_add_pred(nbs);                                                                  \
        uint32_t lo, hi;                                                                                                \
        ring_bits64(cur - incl, lo, hi);        /* state 4l's bits are the uppermost of the lane's window */            \
        const uint32_t tot = __shfl_sync(FULL, incl, 31);                                                               \
        if (tot > cur - floor_bits) { bad = 1; break; }                                                                 \
        const uint32_t w2 = __funnelshift_r(lo, hi, n3), w1 = __funnelshift_r(lo, hi, n23);   /* n23 <= 22 */           \
        const uint32_t w0 = __funnelshift_r(w1, hi >> n23, n1);                                                         \
        a3 = SHD_NEXT(e3, lo, n3);                                                                                      \
        a2 = SHD_NEXT(e2, w2, n2);                                                                                      \
        a1 = SHD_NEXT(e1, w1, n1);                                                                                      \
        a0 = SHD_NEXT(e0, w0, n0);                                                                                      \
        const uint32_t sy = __byte_perm(__byte_perm(e0, e1, 0x0062), __byte_perm(e2, e3, 0x0062), 0x5410);   /* byte 2 of e0..e3 */ \
        STORE;                                                                                                          \
        cur -= tot;                                                                                                     \
    }
    if (out_aligned) {
        uint32_t *const ow = reinterpret_cast<uint32_t *>(out);              // warp uniform; the index stays 32 bits wide
        for (; i0 + 128 <= body; i0 += 128) SHD_ROUND(ow[(i0 >> 2) + (uint32_t)lane] = sy)
    } else {
        uint8_t *ob = out + 4 * lane;
        for (; i0 + 128 <= body; i0 += 128, ob += 128)
            SHD_ROUND(ob[0] = (uint8_t)sy; ob[1] = (uint8_t)(sy >> 8); ob[2] = (uint8_t)(sy >> 16); ob[3] = (uint8_t)(sy >> 24))
    }
#undef SHD_ROUND
    if (!bad && i0 < body) {                                // last partial round
        stage();
        const uint32_t ia = i0 + 4 * lane;
        const uint32_t e0 = lds_u32(a0), e1 = lds_u32(a1), e2 = lds_u32(a2), e3 = lds_u32(a3);
        const uint32_t n0 = ia < body ? (e0 >> 28) : 0u, n1 = ia + 1 < body ? (e1 >> 28) : 0u;
        const uint32_t n2 = ia + 2 < body ? (e2 >> 28) : 0u, n3 = ia + 3 < body ? (e3 >> 28) : 0u;
        const uint32_t n23 = n2 + n3, n123 = n1 + n23, nbs = n0 + n123;
        const uint32_t incl = warp_incl_add_pred(nbs);
        const uint32_t tot = __shfl_sync(FULL, incl, 31);
        if (tot > cur - floor_bits) bad = 1;
        else {
            uint32_t lo, hi;
            ring_bits64(cur - incl, lo, hi);
            const uint64_t w = ((uint64_t)hi << 32) | lo;
            if (ia < body) { out[ia] = (uint8_t)SHD_SYM(e0); a0 = SHD_NEXT(e0, (uint32_t)(w >> n123), n0); }
            if (ia + 1 < body) { out[ia + 1] = (uint8_t)SHD_SYM(e1); a1 = SHD_NEXT(e1, (uint32_t)(w >> n23), n1); }
            if (ia + 2 < body) { out[ia + 2] = (uint8_t)SHD_SYM(e2); a2 = SHD_NEXT(e2, (uint32_t)(w >> n3), n2); }
            if (ia + 3 < body) { out[ia + 3] = (uint8_t)SHD_SYM(e3); a3 = SHD_NEXT(e3, (uint32_t)w, n3); }
            cur -= tot;
        }
    }
#undef SHD_NEXT
#undef SHD_NB
    if (!bad) {                                             // Decoder::finish, fse.rs:383-385: i in [body, bn), state i % 128
        out[body + ((4 * lane - body) & 127)] = (uint8_t)SHD_SYM(lds_u32(a0));
        out[body + ((4 * lane + 1 - body) & 127)] = (uint8_t)SHD_SYM(lds_u32(a1));
        out[body + ((4 * lane + 2 - body) & 127)] = (uint8_t)SHD_SYM(lds_u32(a2));
        out[body + ((4 * lane + 3 - body) & 127)] = (uint8_t)SHD_SYM(lds_u32(a3));
    }
#undef SHD_SYM
    if (pending) {                                          // never leave a copy in flight into memory the next stream reuses
        if (!mbar_wait(wk.bar, wk.par)) tma_ok = 0;
        wk.par ^= 1;
    }
    if (!tma_ok) bad = 1;
    cur -= floor_bits;
    return (bad || cur != 0) ? ST_LENGTH : ST_OK;
}

__device__ __forceinline__ int sh_decode_dispatch(uint32_t R, const uint8_t *pay, uint32_t plen, uint8_t *out, uint32_t bn, uint32_t log2,
                                                  uint32_t row0, uint32_t lanepart, ShDecWarp &wk, int lane)
{
    if (R == 32) return sh_decode_payload_warp<7>(pay, plen, out, bn, log2, row0, lanepart, wk, lane);
    if (R == 16) return sh_decode_payload_warp<6>(pay, plen, out, bn, log2, row0, lanepart, wk, lane);
    return sh_decode_payload_warp<5>(pay, plen, out, bn, log2, row0, lanepart, wk, lane);
}

// global-table mode: every warp of the CTA decodes its own blocks against the CTA's replicated copy of the one table
__global__ void __launch_bounds__(1024) k_decode_sh_global(DecArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const uint32_t log2 = a.g.log2;
    const uint32_t R = a.dec_copies;
    const ShDecLayout lay = sh_dec_layout(log2, R, warps);
    // rows must start on a 128-byte boundary: the lane part of an address is OR-ed into the row address
    uint8_t *smem = smem_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u);
    uint8_t *tabR = smem + lay.tab;
    const uint32_t tab_saddr = (uint32_t)__cvta_generic_to_shared(tabR);
    sh_replicate_dec(a.g.dec_table, log2, R, tabR, tab_saddr, threadIdx.x, blockDim.x);
    uint8_t *mine = smem + lay.ring + (size_t)warp * SH_DEC_RING_BYTES;
    ShDecWarp wk;
    wk.ring = reinterpret_cast<uint32_t *>(mine);
    wk.ring_saddr = (uint32_t)__cvta_generic_to_shared(mine);
    wk.bar = wk.ring_saddr + 1024 + 16;
    wk.par = 0;
    if (lane == 0) mbar_init(wk.bar, 1);
    const uint32_t lanepart = ((uint32_t)lane & (R - 1u)) << 2;
    const uint32_t N = 128;
    __shared__ uint32_t cta_next;
    const uint32_t cta_first = (uint32_t)(((unsigned long long)a.nblocks * blockIdx.x) / gridDim.x);
    const uint32_t cta_last = (uint32_t)(((unsigned long long)a.nblocks * (blockIdx.x + 1)) / gridDim.x);
    if (threadIdx.x == 0) cta_next = cta_first;
    __syncthreads();
    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(&cta_next, 1u);
        b = __shfl_sync(FULL, b, 0);
        if (b >= cta_last) break;
        const size_t off = (size_t)b * a.block_size;
        const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
        uint8_t *out = a.dst + off;
        int st = ST_OK;
        const uint8_t *cs;
        uint32_t clen;
        if (!dec_block_prologue(a, b, bn, N, out, lane, cs, clen, st))       // bad offsets and the short raw tail are handled there
            st = sh_decode_dispatch(R, cs, clen, out, bn, log2, tab_saddr, lanepart, wk, lane);
        __syncwarp();
        if (lane == 0) a.status[b] = st;
    }
}

// ------------------------------------------------------------------------------------------
// Segmented per-block mode (see fse_shared_enc.cuh): one CTA per block; the builder warp parses the header of the
// NEXT block and builds its decode table while the other warps decode the segments of the current one.
// ------------------------------------------------------------------------------------------
struct ShDecMeta { uint32_t log2, consumed; int kind; };   // kind 0: coded; 1 / 2: escape handled by the builder; < 0: error
constexpr uint32_t SH_DEC_BUILD_BYTES = 8192 + 1024 + 1024 + 2048 + 528;   // table u32[2048] | norm | ctr | spread | staged header
struct ShDecBlocksLayout { uint32_t tab, build, meta, ring, total; };
__host__ __device__ inline ShDecBlocksLayout sh_dec_blocks_layout(uint32_t tlmax, uint32_t R, int coder_warps)
{
    ShDecBlocksLayout l;
    l.tab = 0;
    l.build = (1u << tlmax) * 4u * R;
    l.meta = l.build + SH_DEC_BUILD_BYTES;
    l.ring = l.meta + 32;
    l.total = l.ring + SH_DEC_RING_BYTES * (uint32_t)coder_warps + 128;
    return l;
}

__device__ __forceinline__ void sh_build_dec_block(const DecArgs &a, uint32_t b, uint8_t *build, ShDecMeta *meta, int lane)
{
    uint32_t *tab = reinterpret_cast<uint32_t *>(build);
    int32_t *norm = reinterpret_cast<int32_t *>(build + 8192);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(build + 9216);
    uint8_t *spread = build + 10240;
    uint32_t *hw = reinterpret_cast<uint32_t *>(build + 12288);
    const size_t off = (size_t)b * a.block_size;
    const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
    uint8_t *out = a.dst + off;
    const uint32_t S = a.segs_per_block, s0 = b * S;
    const uint32_t nseg = (bn + a.seg_size - 1) / a.seg_size;
    const unsigned long long o0 = a.offsets[s0], o1 = a.offsets[s0 + 1];
    uint32_t log2 = 0, consumed = 0;
    int kind = 0;
    do {
        if (o1 < o0 || o1 > a.comp_bytes || o1 - o0 > 0xffffffffull) { kind = ST_LENGTH; break; }
        const uint8_t *cs = a.comp + o0;
        const uint32_t clen = (uint32_t)(o1 - o0);
        if (clen == 0) { kind = ST_PANIC; break; }
        const uint32_t first = cs[0];
        if ((first & 0x0f) == 0x0f) {                        // raw escape: the whole block
            if (clen != 1 + bn) { kind = ST_LENGTH; break; }
            for (uint32_t i = lane; i < bn; i += 32) out[i] = cs[1 + i];
            kind = 1;
            break;
        }
        if ((first & 0x0f) == 0x0e) {                        // run escape
            if (clen != 2) { kind = ST_LENGTH; break; }
            const uint8_t v = cs[1];
            for (uint32_t i = lane; i < bn; i += 32) out[i] = v;
            kind = 2;
            break;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) norm[k * 32 + lane] = 0;
        __syncwarp();
        uint32_t table_len = 0;
        const int rc = warp_ncount_read(cs, clen, hw, norm, lane, log2, table_len, consumed);
        if (rc < 0) { kind = rc; break; }
        if (log2 > a.tlmax || log2 > 11) { kind = ST_UNSUPPORTED; break; }
        warp_spread(norm, log2, table_len, spread, ctr, reinterpret_cast<uint16_t *>(tab), lane);
        warp_build_decode(norm, log2, table_len, spread, ctr, tab, lane);
    } while (0);
    __syncwarp();
    if (kind != 0)
        for (uint32_t k = lane; k < nseg; k += 32) a.status[s0 + k] = kind;
    if (lane == 0) { meta->log2 = log2; meta->consumed = consumed; meta->kind = kind; }
    __syncwarp();
}

__global__ void __launch_bounds__(544, 2) k_decode_sh_blocks(DecArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int coders = warps - 1;
    const uint32_t R = a.dec_copies;
    const ShDecBlocksLayout lay = sh_dec_blocks_layout(a.tlmax, R, coders);
    uint8_t *smem = smem_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u);
    uint8_t *tabR = smem + lay.tab, *build = smem + lay.build;
    ShDecMeta *meta = reinterpret_cast<ShDecMeta *>(smem + lay.meta);
    const uint32_t tab_saddr = (uint32_t)__cvta_generic_to_shared(tabR);
    uint8_t *mine = smem + lay.ring + (size_t)(warp < coders ? warp : 0) * SH_DEC_RING_BYTES;
    ShDecWarp wk;
    wk.ring = reinterpret_cast<uint32_t *>(mine);
    wk.ring_saddr = (uint32_t)__cvta_generic_to_shared(mine);
    wk.bar = wk.ring_saddr + 1024 + 16;
    wk.par = 0;
    const bool builder = warp == coders;
    if (!builder && lane == 0) mbar_init(wk.bar, 1);
    const uint32_t lanepart = ((uint32_t)lane & (R - 1u)) << 2;
    const uint32_t first = (uint32_t)(((unsigned long long)a.nblocks * blockIdx.x) / gridDim.x);
    const uint32_t last = (uint32_t)(((unsigned long long)a.nblocks * (blockIdx.x + 1)) / gridDim.x);
    if (builder && first < last) sh_build_dec_block(a, first, build, meta, lane);
    __syncthreads();
    for (uint32_t b = first; b < last; b++) {
        const uint32_t log2 = meta->log2, consumed = meta->consumed;
        const int kind = meta->kind;
        if (kind == 0) sh_replicate_dec(reinterpret_cast<const uint32_t *>(build), log2, R, tabR, tab_saddr, threadIdx.x, blockDim.x);
        __syncthreads();
        if (builder) {
            if (b + 1 < last) sh_build_dec_block(a, b + 1, build, meta, lane);
        } else if (kind == 0) {
            const size_t off = (size_t)b * a.block_size;
            const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
            const uint32_t nseg = (bn + a.seg_size - 1) / a.seg_size;
            for (uint32_t k = warp; k < nseg; k += coders) {
                const uint32_t slot = b * a.segs_per_block + k;
                const uint32_t so = k * a.seg_size, sn = min(a.seg_size, bn - so);
                uint8_t *out = a.dst + off + so;
                const unsigned long long o0 = a.offsets[slot], o1 = a.offsets[slot + 1];
                int st;
                const uint32_t skip = k == 0 ? consumed : 0u;
                if (o1 < o0 || o1 > a.comp_bytes || o1 - o0 > 0xffffffffull || o1 - o0 < skip) st = ST_LENGTH;
                else {
                    const uint8_t *cs = a.comp + o0 + skip;
                    const uint32_t clen = (uint32_t)(o1 - o0) - skip;
                    if (sn < 128) {
                        if (clen != sn) st = ST_LENGTH;
                        else {
                            for (uint32_t i = lane; i < sn; i += 32) out[i] = cs[i];
                            st = 1;
                        }
                    } else st = sh_decode_dispatch(R, cs, clen, out, sn, log2, tab_saddr, lanepart, wk, lane);
                }
                __syncwarp();
                if (lane == 0) a.status[slot] = st;
            }
        }
        __syncthreads();
    }
}

}  // namespace fsed
