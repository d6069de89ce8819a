// fse_decode64c.cuh -- 64-state decode with a compact per-warp footprint (table_log <= 12).
//
// The decoder is a serial chain per warp (table lookup -> prefix of bit counts -> bit fetch -> next
// state), so throughput is set by how many warps an SM can hold, and that is set by shared memory.
// The reference's DecodeTransform {u16 new_state, u8 symbol, u8 num_bits} (fse.rs:260-265) is split:
//   tab16[cell] = new_state | num_bits << 12      (on the critical path)
//   sym[cell]   = symbol                          (the spread itself, off the critical path)
// and the build scratch is aliased with the payload ring: 3*size + 2 KiB per warp = 8 KiB at
// table_log 11, i.e. 28 warps (blocks in flight) per SM instead of 14.
#pragma once
#include "fse_kernels64.cuh"
#ifndef FSE_DEC_SCAN
#define FSE_DEC_SCAN 0   /* 1 = ballot+popc prefix: measured slower (0.59 vs 0.49 ms on c2) */
#endif

namespace fsed {

struct Dec64cLayout {
    uint32_t tab;      // uint16[size]
    uint32_t sym;      // uint8[size]
    uint32_t scratch;  // build: norm i32[256], then the counters u32[256] in the same array; decode: ring u32[258]
    uint32_t total;
};
__host__ __device__ inline Dec64cLayout dec64c_layout(uint32_t tlmax)
{
    Dec64cLayout l;
    uint32_t size = 1u << tlmax;
    l.tab = 0;
    l.sym = size * 2;
    l.scratch = size * 3;
    l.total = size * 3 + 1040 + 16;     // + one mbarrier per warp (fse_decode128c.cuh); 16 warps per CTA at table_log 11
    return l;
}

// Inclusive warp prefix sum with the classic predicated add: shfl.up returns "source lane valid" as a
// predicate, so each step is SHFL + @p IADD (no select, nothing on the ALU pipe).  A radix-4 form (three dependent
// shuffle levels, distances 1 2 3 | 4 8 12 | 16, 7 + 7 instructions) was measured in round 2: c4 8.46 against 8.54 ms,
// c5 6.87 against 6.94, c2 0.387 against 0.378: the scan's latency is not what the decoders wait for.
__device__ __forceinline__ uint32_t warp_incl_add_pred(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
        asm volatile("{ .reg .u32 t; .reg .pred p; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, t; }"
                     : "+r"(v) : "r"(d));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32_4(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(v) : "r"(saddr));
    return v;
}

// DecodeTable::update, fse.rs:328-337, compact form
template <bool INIT = true>     // false: ctr already holds the counters (warp_spread<true>)
__device__ __forceinline__ void warp_build_decode16(const int32_t *norm, uint32_t log2, uint32_t table_len,
                                                    const uint8_t *spread, uint32_t *ctr, uint16_t *table, int lane)
{
    const uint32_t size = 1u << log2;
    if (INIT) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int i = lane * 8 + k;
            int32_t x = (i < (int)table_len) ? norm[i] : 0;
            ctr[i] = (x < 0) ? 1u : (uint32_t)x;
        }
        __syncwarp();
    }
    for (uint32_t c0 = 0; c0 < size; c0 += 32) {
        uint32_t cell = c0 + lane;
        uint32_t s = spread[cell];
        uint32_t m = __match_any_sync(FULL, s);
        uint32_t r = __popc(m & lt_mask(lane));
        uint32_t base = ctr[s];
        __syncwarp();
        if (r == 0) ctr[s] = base + __popc(m);
        uint32_t next = base + r;
        uint32_t nb = log2 - ilog2u(next | (next == 0));
        table[cell] = (uint16_t)((((next << nb) - size) & 0xfffu) | (nb << 12));
        __syncwarp();
    }
}

__global__ void __launch_bounds__(512, 2) k_decode64c_blocks(DecArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const Dec64cLayout lay = dec64c_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint16_t *tab = reinterpret_cast<uint16_t *>(my + lay.tab);
    uint8_t *sym = my + lay.sym;
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.scratch);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(my + lay.scratch);          // norm's own array (warp_spread<true>)
    uint32_t *ring = reinterpret_cast<uint32_t *>(my + lay.scratch);
    const uint32_t N = 64;

    uint32_t glog2 = 0;
    if (a.global_mode) {
        glog2 = a.g.log2;
        for (uint32_t i = lane; i < (1u << glog2); i += 32) {
            uint32_t e = a.g.dec_table[i];
            tab[i] = (uint16_t)((e & 0xfffu) | ((e >> 24) << 12));
            sym[i] = (uint8_t)(e >> 16);
        }
        __syncwarp();
    }

    for (uint32_t b = blockIdx.x * wpc + warp; b < a.nblocks; b += gridDim.x * wpc) {
        const size_t off = (size_t)b * a.block_size;
        const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
        uint8_t *out = a.dst + off;
        int st = ST_OK;
        const uint8_t *cs;
        uint32_t clen;
        if (dec_block_prologue(a, b, bn, N, out, lane, cs, clen, st)) {      // bad offsets, raw tail, escape blocks
            if (lane == 0) { a.status[b] = st; if (a.exhaust) a.out_len[b] = 0; }
            continue;
        }
        uint32_t log2 = glog2, consumed = 0;
        __syncwarp();
        if (!a.global_mode) {
#pragma unroll
            for (int k = 0; k < 8; k++) norm[k * 32 + lane] = 0;
            __syncwarp();
            uint32_t table_len = 0;
            const int rc = warp_ncount_read(cs, clen, reinterpret_cast<uint32_t *>(tab), norm, lane, log2, table_len, consumed);
            if (rc < 0) { if (lane == 0) a.status[b] = rc; continue; }
            if (log2 > a.tlmax || log2 > 12) { if (lane == 0) a.status[b] = ST_UNSUPPORTED; continue; }
            warp_spread<true>(norm, log2, table_len, sym, ctr, tab, lane);     // sym is the spread; tab doubles as posmap
            warp_build_decode16<false>(norm, log2, table_len, sym, ctr, tab, lane);
        }
        if (bn < N) { if (lane == 0) a.status[b] = ST_LENGTH; continue; }
        const uint8_t *pay = cs + consumed;
        const uint32_t plen = clen - consumed;
        if (plen == 0 || pay[plen - 1] == 0) { if (lane == 0) a.status[b] = ST_NO_MARKER; continue; }
        const uint32_t bias = (uint32_t)((uintptr_t)pay & 3);
        const uint32_t *origin = reinterpret_cast<const uint32_t *>(pay - bias);
        uint32_t cur = (plen - 1) * 8 + ilog2u(pay[plen - 1]) + 8 * bias;
        const uint32_t floor_bits = 8 * bias;
        if (cur - floor_bits < N * log2) { if (lane == 0) a.status[b] = ST_LENGTH; continue; }
        const uint32_t topq = cur >> 5;
        uint32_t lowq = (topq & ~127u) >= 128 ? (topq & ~127u) - 128 : 0;
        __syncwarp();                                       // the build scratch becomes the ring
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t w = lowq + lane + 32 * k;
            if (w <= topq) {
                uint32_t x = __ldg(origin + w);
                ring[w & 255] = x;
                if ((w & 255) == 0) ring[256] = x;          // mirror: ring[256] == ring[0]
            }
        }
        uint32_t pre[4];
#pragma unroll
        for (int k = 0; k < 4; k++) pre[k] = (lowq >= 128) ? __ldg(origin + lowq - 128 + lane + 32 * k) : 0u;
        __syncwarp();
        const uint32_t ring_saddr = (uint32_t)__cvta_generic_to_shared(ring);
        auto ring_bits = [&](uint32_t q, uint32_t nb) -> uint32_t {
            uint32_t a = ring_saddr + ((q >> 3) & 0x3fcu);
            return __funnelshift_r(lds_u32(a), lds_u32_4(a), q & 31) & ~(0xffffffffu << nb);
        };
        uint32_t st0, st1;                                  // Decoder::new, fse.rs:349-352
        {
            uint32_t w = ring_bits(cur - (2 * lane + 2) * log2, 2 * log2);
            st0 = w >> log2;
            st1 = w & ((1u << log2) - 1u);
        }
        cur -= N * log2;
        const uint32_t body = bn - N;
        const bool out_aligned = (((uintptr_t)out) & 1) == 0;
        bool bad = false;
        uint32_t i0 = 0;
        for (; i0 + 64 <= body; i0 += 64) {
            if ((cur >> 5) < lowq + 28 && lowq) {           // a round takes at most 24 words: refill the ring
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; k++) ring[(lowq - 128 + lane + 32 * k) & 255] = pre[k];
                if (lane == 0 && ((lowq - 128) & 255) == 0) ring[256] = pre[0];
                lowq -= 128;
#pragma unroll
                for (int k = 0; k < 4; k++) pre[k] = (lowq >= 128) ? __ldg(origin + lowq - 128 + lane + 32 * k) : 0u;
                __syncwarp();
            }
            uint32_t e0 = tab[st0], e1 = tab[st1];          // fse.rs:363-373, two independent chains
            uint32_t y0 = sym[st0], y1 = sym[st1];
            uint32_t nb0 = e0 >> 12, nb1 = e1 >> 12;
            uint32_t nbs = nb0 + nb1;
#if FSE_DEC_SCAN == 1
            uint32_t incl = warp_incl_add5(nbs, lane);      // ballots + popc: no shared-memory crossbar traffic
#else
            uint32_t incl = warp_incl_add_pred(nbs);
#endif
            uint32_t w = ring_bits(cur - incl, nbs);        // state 2l's bits are the upper part
            uint32_t tot = __shfl_sync(FULL, incl, 31);
            if (tot > cur - floor_bits) { bad = true; break; }
            st0 = (e0 & 0xfffu) + (w >> nb1);
            st1 = (e1 & 0xfffu) + (w & ((1u << nb1) - 1u));
            uint32_t sy = y0 | (y1 << 8);
            if (out_aligned) *reinterpret_cast<uint16_t *>(out + i0 + 2 * lane) = (uint16_t)sy;
            else { out[i0 + 2 * lane] = (uint8_t)y0; out[i0 + 2 * lane + 1] = (uint8_t)y1; }
            cur -= tot;
        }
        if (!bad && i0 < body) {                            // last partial round
            if ((cur >> 5) < lowq + 28 && lowq) {
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; k++) ring[(lowq - 128 + lane + 32 * k) & 255] = pre[k];
                if (lane == 0 && ((lowq - 128) & 255) == 0) ring[256] = pre[0];
                lowq -= 128;
                __syncwarp();
            }
            uint32_t ia = i0 + 2 * lane, ib = ia + 1;
            uint32_t e0 = tab[st0], e1 = tab[st1];
            uint32_t nb0 = ia < body ? (e0 >> 12) : 0u, nb1 = ib < body ? (e1 >> 12) : 0u;
            uint32_t nbs = nb0 + nb1;
            uint32_t incl = warp_incl_add5(nbs, lane);
            uint32_t tot = __shfl_sync(FULL, incl, 31);
            if (tot > cur - floor_bits) bad = true;
            else {
                uint32_t w = ring_bits(cur - incl, nbs);
                if (ia < body) { out[ia] = sym[st0]; st0 = (e0 & 0xfffu) + (w >> nb1); }
                if (ib < body) { out[ib] = sym[st1]; st1 = (e1 & 0xfffu) + (w & ((1u << nb1) - 1u)); }
                cur -= tot;
            }
        }
        if (!bad) {                                         // Decoder::finish, fse.rs:383-385
            uint32_t ia = body + ((2 * lane - body) & 63), ib = body + ((2 * lane + 1 - body) & 63);
            out[ia] = sym[st0];
            out[ib] = sym[st1];
        }
        cur -= floor_bits;
        if (bad || cur != 0) st = ST_LENGTH;
        if (lane == 0) a.status[b] = st;
    }
}

}  // namespace fsed
