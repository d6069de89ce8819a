// fse_decode64w.cuh -- 64-state decode with full DecodeTransform entries (fse.rs:260-265: one 32-bit
// load per state gives new_state, symbol and num_bits) in a 4*size + 2 KiB footprint: the spread lives
// in the last quarter of the table region and is consumed in place by the table build (entry c only
// overwrites spread cells <= c), the build scratch is aliased with the payload ring.  10 KiB per warp at
// table_log 11 = 22 warps per SM; against the compact layout (fse_decode64c.cuh, 28 warps) it trades
// warps for two fewer bank-conflicted loads per round.
#pragma once
#include "fse_decode64c.cuh"

namespace fsed {

struct Dec64wLayout {
    uint32_t tab;      // uint32[size]
    uint32_t sym;      // uint8[size], aliased with tab's last quarter
    uint32_t scratch;  // build: norm i32[256] | ctr u32[256]; decode: ring u32[257]
    uint32_t total;
};
__host__ __device__ inline Dec64wLayout dec64w_layout(uint32_t tlmax)
{
    Dec64wLayout l;
    uint32_t size = 1u << tlmax;
    l.tab = 0;
    l.sym = size * 3;
    l.scratch = size * 4;
    l.total = size * 4 + 2048;
    return l;
}

__global__ void __launch_bounds__(512, 2) k_decode64w_blocks(DecArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const Dec64wLayout lay = dec64w_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint32_t *tab = reinterpret_cast<uint32_t *>(my + lay.tab);
    uint8_t *sym = my + lay.sym;                            // the spread, consumed in place by the table build
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.scratch);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(my + lay.scratch + 1024);
    uint32_t *ring = reinterpret_cast<uint32_t *>(my + lay.scratch);
    const uint32_t N = 64;

    uint32_t glog2 = 0;
    if (a.global_mode) {
        glog2 = a.g.log2;
        for (uint32_t i = lane; i < (1u << glog2); i += 32) {
            tab[i] = a.g.dec_table[i];
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
            int rc = 0;
            if (lane == 0) rc = ncount_read_serial(cs, clen, norm, log2, table_len, consumed);
            rc = __shfl_sync(FULL, rc, 0);
            log2 = __shfl_sync(FULL, log2, 0);
            table_len = __shfl_sync(FULL, table_len, 0);
            consumed = __shfl_sync(FULL, consumed, 0);
            __syncwarp();
            if (rc < 0) { if (lane == 0) a.status[b] = rc; continue; }
            if (log2 > a.tlmax || log2 > 13) { if (lane == 0) a.status[b] = ST_UNSUPPORTED; continue; }
            warp_spread(norm, log2, table_len, sym, ctr, reinterpret_cast<uint16_t *>(tab), lane);   // posmap: first half of tab
            warp_build_decode(norm, log2, table_len, sym, ctr, tab, lane);   // entry c overwrites spread cells <= c only
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
            uint32_t y0 = (e0 >> 16) & 0xffu, y1 = (e1 >> 16) & 0xffu;
            uint32_t nb0 = e0 >> 24, nb1 = e1 >> 24;
            uint32_t nbs = nb0 + nb1;
            uint32_t incl = warp_incl_add_pred(nbs);
            uint32_t w = ring_bits(cur - incl, nbs);        // state 2l's bits are the upper part
            uint32_t tot = __shfl_sync(FULL, incl, 31);
            if (tot > cur - floor_bits) { bad = true; break; }
            st0 = (e0 & 0xffffu) + (w >> nb1);
            st1 = (e1 & 0xffffu) + (w & ((1u << nb1) - 1u));
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
            uint32_t nb0 = ia < body ? (e0 >> 24) : 0u, nb1 = ib < body ? (e1 >> 24) : 0u;
            uint32_t nbs = nb0 + nb1;
            uint32_t incl = warp_incl_add5(nbs, lane);
            uint32_t tot = __shfl_sync(FULL, incl, 31);
            if (tot > cur - floor_bits) bad = true;
            else {
                uint32_t w = ring_bits(cur - incl, nbs);
                if (ia < body) { out[ia] = (uint8_t)(e0 >> 16); st0 = (e0 & 0xffffu) + (w >> nb1); }
                if (ib < body) { out[ib] = (uint8_t)(e1 >> 16); st1 = (e1 & 0xffffu) + (w & ((1u << nb1) - 1u)); }
                cur -= tot;
            }
        }
        if (!bad) {                                         // Decoder::finish, fse.rs:383-385
            uint32_t ia = body + ((2 * lane - body) & 63), ib = body + ((2 * lane + 1 - body) & 63);
            out[ia] = (uint8_t)(tab[st0] >> 16);
            out[ib] = (uint8_t)(tab[st1] >> 16);
        }
        cur -= floor_bits;
        if (bad || cur != 0) st = ST_LENGTH;
        if (lane == 0) a.status[b] = st;
    }
}

}  // namespace fsed
