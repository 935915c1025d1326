// lz4_kernels.cu — hand-written sm_100a kernels for the LZ4 side of the zstdmt hot path.
//
// Replaces the arithmetic the reference reaches through
//   LZ4F_compressFrame   (/root/reference/lib/lz4-mt_compress.c:280-283)
//   LZ4F_decompress      (/root/reference/lib/lz4-mt_decompress.c:349-351)
// plus the 12-byte skippable container write (lib/lz4-mt_compress.c:293-298).
//
// Kernels
//   lz4_compress_blocks_kernel  one CTA per 64 KiB LZ4F block; block staged in SMEM by the TMA
//                               unit (cp.async.bulk); round-synchronous hash candidates,
//                               speculative-chain parallel greedy parse, scan-based emission.
//   xxh32_kernel                4 lanes per chunk (the 4 XXH32 accumulators), 8 chunks per warp.
//   lz4_frame_sizes_kernel / lz4_frame_pack_kernel
//                               frame size per chunk -> exclusive scan -> compaction of the block
//                               payloads + LZ4F header/end-mark/checksum + skippable header.
//   lz4_decode_frames_kernel    one warp per frame, warp-lockstep sequence decode.
//
// The compressor's match/parse rule is deterministic and restated on the CPU in
// oracle/lz4_oracle.c:orc_lz4_block_compress_b200 (tests compare bit-exact).
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "zmt_dev.h"

// bytes of chunk c: explicit per-chunk sizes (host pipeline: any fn_read may come back short)
// or derived from a contiguous input of in_bytes cut every chunk_size bytes.
__device__ __forceinline__ uint64_t zmt_chunk_len(const uint32_t* __restrict__ cb, uint32_t c, uint64_t in_bytes, uint32_t chunk_size)
{
    if (cb) return cb[c];
    const uint64_t base = (uint64_t)c * chunk_size;
    if (in_bytes <= base) return 0;
    return (in_bytes - base) < chunk_size ? (in_bytes - base) : chunk_size;
}

// A CTA barrier preceded by a warp reconvergence: the data-dependent phases leave warps split into
// independently scheduled groups, and an aligned BAR.SYNC + the uniform-datapath branches ptxas places
// after it (BRA.U) must only ever be executed by whole warps (else: 'illegal instruction' on sm_100).
#define CTA_SYNC() do { __syncwarp(); __syncthreads(); } while (0)

// ============================================================================ compressor
#define LZ4_BLK      65536u
#define C_NT         256u          // threads per CTA
#define C_TILE       4096u         // positions parsed per tile
#define C_SEG        16u           // C_TILE / C_NT : positions owned by one speculative chain
#define C_ROUND      1024u         // hash-table update granularity
#define C_HASHLOG    12
#define C_MAXSEQ     1024u         // max sequences per tile (min advance 4)
#define C_LONGLIT    32u           // literal runs longer than this are copied cooperatively
#define C_END        0xFFFFu       // link: chain leaves the tile

struct __align__(16) CompressSmem {
    uint8_t  in[LZ4_BLK + 32];            // block bytes + zero pad
    uint32_t tab[1 << C_HASHLOG];         // hash -> 1 + position
    uint16_t off[C_TILE];                 // per tile position: match offset (0 = none)
    uint16_t len[C_TILE];                 // per visited match start: match length
    uint32_t M[C_TILE / 32];              // has-match bits
    uint32_t V[C_TILE / 32];              // visited-by-own-chain bits
    uint32_t Sel[C_TILE / 32];            // selected (true greedy chain) bits
    uint32_t xfree[C_NT];                 // own-walk exit (free position, absolute)
    uint32_t mpos[C_NT];                  // merge position / tile exit (absolute)
    uint32_t min_[C_NT];                  // entry free position of a reachable chain
    uint16_t link[C_NT];                  // chain this chain merges into (or C_END / dead)
    uint16_t jump[C_NT];
    uint8_t  reach[C_NT];
    uint16_t seqpos[C_MAXSEQ];            // tile-relative start of the r-th selected sequence
    uint32_t longl[3 * (C_TILE / C_LONGLIT + 1)];
    uint32_t scanws[40];
    uint32_t nlong;
    uint32_t e_next;                      // chain entry for the next tile
    uint64_t mbar;
};

__device__ __forceinline__ uint32_t c_match_len(const uint8_t* s, uint32_t q, uint32_t c, uint32_t limit)
{
    uint32_t L = 4;
    while (q + L + 4 <= limit) {
        uint32_t x = lds32u(s, q + L) ^ lds32u(s, c + L);
        if (x) return L + ((__ffs(x) - 1) >> 3);
        L += 4;
    }
    while (q + L < limit && s[q + L] == s[c + L]) L++;
    return L;
}

// Walk one speculative chain.  MODE 0: own segment only (sets V, caches len, returns exit in xfree)
//                              MODE 1: continuation until merge/end (records link/mpos)
//                              MODE 2: re-walk of a reachable chain (marks Sel) — same steps as 0+1.
template <int MODE>
__device__ __forceinline__ void c_walk(CompressSmem& S, uint32_t k, uint32_t p, uint32_t t0, uint32_t limit)
{
    // Single-exit loop (no returns from inside): divergent lanes leave through one reconvergence
    // point, so the warp is whole again before the CTA barrier that follows every call.
    const uint32_t t1 = t0 + C_TILE;
    uint32_t lk = 0xFFFFFFFFu, mp = 0;          // result: link / merge position (MODE 1), exit (MODE 0)
    bool done = false;
    while (!done) {
        if (p >= t1) { lk = C_END; mp = p; done = true; continue; }
        const uint32_t rel = p - t0, j = rel / C_SEG;
        if (MODE == 0 && j != k) { mp = p; done = true; continue; }
        if (MODE != 0 && j != k && (rel & (C_SEG - 1)) == 0) { lk = j; mp = p; done = true; continue; }
        uint32_t bits = (S.M[rel >> 5] >> (rel & 16)) & 0xFFFFu;     // this segment's 16 bits
        bits &= 0xFFFFu << (rel & 15);
        if (!bits) {                                                   // segment exhausted -> free position at its end
            const uint32_t nx = t0 + (j + 1) * C_SEG;
            if (MODE == 0) { mp = nx; done = true; }
            else if (j + 1 == C_NT) { lk = C_END; mp = t1; done = true; }
            else if (j != k) { lk = j + 1; mp = nx; done = true; }
            else p = nx;                                               // MODE 2 inside own segment
            continue;
        }
        const uint32_t qr = (rel & ~15u) + (__ffs(bits) - 1), q = t0 + qr;
        if (MODE != 0 && j != k && ((S.V[qr >> 5] >> (qr & 31)) & 1)) { lk = j; mp = q; done = true; continue; }
        uint32_t L;
        if (MODE == 2) { atomicOr(&S.Sel[qr >> 5], 1u << (qr & 31)); L = S.len[qr]; }
        else {
            L = c_match_len(S.in, q, q - S.off[qr], limit);
            S.len[qr] = (uint16_t)L;
            if (MODE == 0) atomicOr(&S.V[qr >> 5], 1u << (qr & 31));
        }
        p = q + L;
    }
    if (MODE == 0) S.xfree[k] = mp;
    if (MODE == 1) { S.link[k] = (uint16_t)lk; S.mpos[k] = mp; }
}

__device__ __forceinline__ uint32_t c_seq_size(uint32_t lit, uint32_t L)
{
    uint32_t s = 1 + lit + 2;
    if (lit >= 15) s += 1 + (lit - 15) / 255;
    if (L - 4 >= 15) s += 1 + (L - 19) / 255;
    return s;
}

#ifndef ZMT_DBG_OCC
#define ZMT_DBG_OCC 2
#endif
__global__ void __launch_bounds__(C_NT, ZMT_DBG_OCC)
lz4_compress_blocks_kernel(const uint8_t* __restrict__ in, uint64_t in_bytes, uint32_t chunk_size, const uint32_t* __restrict__ chunk_bytes,
                           uint32_t bpc, uint8_t* __restrict__ tmp, uint32_t* __restrict__ blk_csize, uint32_t nblocks, uint32_t flags)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CompressSmem& S = *reinterpret_cast<CompressSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (uint32_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const uint32_t chunk = blk / bpc, bic = blk % bpc;
        const uint64_t cbase = (uint64_t)chunk * chunk_size;
        const uint64_t cbytes = zmt_chunk_len(chunk_bytes, chunk, in_bytes, chunk_size);
        const uint64_t boff = (uint64_t)bic * LZ4_BLK;
        uint32_t n = 0;
        if (boff < cbytes) n = (uint32_t)((cbytes - boff) < LZ4_BLK ? (cbytes - boff) : LZ4_BLK);
        if (n == 0) { if (tid == 0) blk_csize[blk] = 0; continue; }
        const uint8_t* src = in + cbase + boff;
        uint8_t* dst = tmp + (uint64_t)blk * ZMT_LZ4_TMP_STRIDE;

        // ---- stage the block into shared memory (TMA bulk copy when 16-byte aligned)
        CTA_SYNC();                                  // previous block fully consumed
        const uint32_t nb16 = ((((uintptr_t)src & 15) == 0) && !(flags & 1)) ? (n & ~15u) : 0;
        if (tid == 0) {
            mbar_init(&S.mbar, 1);
            S.nlong = 0;
        }
        CTA_SYNC();
        if (tid == 0 && nb16) {
            mbar_expect_tx(&S.mbar, nb16);
            bulk_g2s(S.in, src, nb16, &S.mbar);
        }
        for (uint32_t i = nb16 + tid; i < n; i += C_NT) S.in[i] = src[i];
        for (uint32_t i = n + tid; i < ((n + 3) & ~3u) + 32 && i < LZ4_BLK + 32; i += C_NT) S.in[i] = 0;
        for (uint32_t i = tid; i < (1u << C_HASHLOG); i += C_NT) S.tab[i] = 0;
        // one thread observes the TMA completion; the CTA barrier publishes the staged bytes to everyone
        if (tid == 0 && nb16) { mbar_wait(&S.mbar, 0); asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S.mbar))); }
        CTA_SYNC();

        const uint32_t limit = n - 5;                     // matches end at or before n-5 (n >= 13 whenever a match exists)
        uint32_t e = 0, anchor = 0, out_pos = 0;          // CTA-uniform parse state

        for (uint32_t t0 = 0; t0 < n; t0 += C_TILE) {
            // ---------------- phase 1: candidates (4 rounds of 1024 positions)
            if (tid < C_TILE / 32) { S.V[tid] = 0; S.Sel[tid] = 0; }
            uint32_t anyM = 0;
#pragma unroll 1
            for (uint32_t r = 0; r < C_TILE / C_ROUND; r++) {
                uint32_t hreg[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    const uint32_t rel = r * C_ROUND + k * C_NT + tid, i = t0 + rel;
                    const bool ok = (i + 12 <= n);
                    const uint32_t v = lds32u(S.in, i);
                    const uint32_t h = (v * 2654435761u) >> (32 - C_HASHLOG);
                    hreg[k] = ok ? h : 0xFFFFFFFFu;
                    uint32_t o = 0;
#ifdef ZMT_DBG_NOMATCH
                    uint32_t same = 0;
                    for (int l = 0; l < 32; l++) { const uint32_t ov = __shfl_sync(ZMT_FULL_MASK, v, l); if (ov == v && (uint32_t)l < lane) same |= 1u << l; }
#else
                    const uint32_t same = __match_any_sync(ZMT_FULL_MASK, v) & ((1u << lane) - 1);
#endif
                    if (ok) {
                        if (same) o = lane - (31 - __clz(same));
                        else {
                            const uint32_t t = S.tab[h];
                            if (t && lds32u(S.in, t - 1) == v) o = i - (t - 1);
                        }
                    }
                    S.off[rel] = (uint16_t)o;
                    const uint32_t mw = __ballot_sync(ZMT_FULL_MASK, o != 0);
                    if (lane == 0) S.M[rel >> 5] = mw;
                    anyM |= mw;
                }
                CTA_SYNC();
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (hreg[k] != 0xFFFFFFFFu) atomicMax(&S.tab[hreg[k]], t0 + r * C_ROUND + k * C_NT + tid + 1);
                CTA_SYNC();
            }
            const uint32_t t1 = t0 + C_TILE;
#ifdef ZMT_DBG_NOOR
            if (tid == 0) S.e_next = 0;
            CTA_SYNC();
            if (anyM && lane == 0) atomicOr(&S.e_next, 1u);
            CTA_SYNC();
            const int tile_has_match = S.e_next != 0;
            CTA_SYNC();
#else
            __syncwarp();
            const int tile_has_match = __syncthreads_or(anyM != 0);
#endif
            if (!tile_has_match || e >= t1) { if (e < t1) e = t1; continue; }   // nothing to parse in this tile

            if (flags & 2) continue;
            // ---------------- phase 2: speculative chains (own segment, then continuation)
            const uint32_t k0 = (e - t0) / C_SEG;
            const uint32_t seg0 = t0 + tid * C_SEG;
            const bool alive = tid >= k0;
            if (alive) c_walk<0>(S, tid, tid == k0 ? e : seg0, t0, limit);
            else S.link[tid] = (uint16_t)tid;             // dead: self link, never reached
            CTA_SYNC();
            if (alive) c_walk<1>(S, tid, S.xfree[tid], t0, limit);
            CTA_SYNC();
            if (flags & 4) continue;
            // ---------------- phase 3: reachability from k0 by pointer doubling
            {
                uint32_t lk = S.link[tid];
                S.jump[tid] = (uint16_t)(lk == C_END ? tid : lk);
                S.reach[tid] = (tid == k0);
                CTA_SYNC();
#pragma unroll 1
                for (int r = 0; r < 8; r++) {
                    const uint32_t j = S.jump[tid];
                    const uint32_t rk = S.reach[tid];
                    const uint32_t jj = S.jump[j];
                    CTA_SYNC();
                    if (rk) S.reach[j] = 1;
                    S.jump[tid] = (uint16_t)jj;
                    CTA_SYNC();
                }
                if (tid == k0) S.min_[tid] = e;
                if (S.reach[tid]) {
                    if (lk == C_END) S.e_next = S.mpos[tid];
                    else S.min_[lk] = S.mpos[tid];
                }
                CTA_SYNC();
            }
            if (flags & 8) continue;
            // ---------------- phase 4: mark the true chain
            if (S.reach[tid]) c_walk<2>(S, tid, S.min_[tid], t0, limit);
            CTA_SYNC();
            e = S.e_next;

            if (flags & 16) continue;
            // ---------------- phase 5: emit the selected sequences
            uint32_t nseq;
            {
                uint32_t w = tid < C_TILE / 32 ? S.Sel[tid] : 0;
                uint32_t base = block_exscan(__popc(w), S.scanws, &nseq);
                while (w) { uint32_t b = __ffs(w) - 1; w &= w - 1; S.seqpos[base++] = (uint16_t)(tid * 32 + b); }
            }
            CTA_SYNC();
            if (nseq == 0) continue;
            if (flags & 32) { const uint32_t lq = S.seqpos[nseq - 1]; anchor = t0 + lq + S.len[lq]; continue; }
            uint32_t lit4[4], len4[4], off4[4], pe4[4], sz = 0, cnt = 0;
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t r = tid * 4 + k;
                if (r < nseq) {
                    const uint32_t qr = S.seqpos[r];
                    uint32_t pe = anchor;
                    if (r) { const uint32_t pr = S.seqpos[r - 1]; pe = t0 + pr + S.len[pr]; }
                    pe4[k] = pe; lit4[k] = t0 + qr - pe; len4[k] = S.len[qr]; off4[k] = S.off[qr];
                    sz += c_seq_size(lit4[k], len4[k]); cnt++;
                }
            }
            uint32_t total;
            uint32_t o = out_pos + block_exscan(sz, S.scanws, &total);
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                if (k >= cnt) break;
                uint8_t* op = dst + o;
                const uint32_t lit = lit4[k], ml = len4[k] - 4;
                *op++ = (uint8_t)(((lit >= 15 ? 15u : lit) << 4) | (ml >= 15 ? 15u : ml));
                if (lit >= 15) { uint32_t x = lit - 15; while (x >= 255) { *op++ = 255; x -= 255; } *op++ = (uint8_t)x; }
                if (lit <= C_LONGLIT) { for (uint32_t i = 0; i < lit; i++) op[i] = S.in[pe4[k] + i]; }
                else { const uint32_t s = atomicAdd(&S.nlong, 1u); S.longl[3 * s] = pe4[k]; S.longl[3 * s + 1] = (uint32_t)(op - dst); S.longl[3 * s + 2] = lit; }
                op += lit;
                *op++ = (uint8_t)off4[k]; *op++ = (uint8_t)(off4[k] >> 8);
                if (ml >= 15) { uint32_t x = ml - 15; while (x >= 255) { *op++ = 255; x -= 255; } *op++ = (uint8_t)x; }
                o += c_seq_size(lit, len4[k]);
            }
            CTA_SYNC();
            {   // cooperative copies of long literal runs: one warp per run
                const uint32_t nl = S.nlong;
                for (uint32_t s = wid; s < nl; s += C_NT / 32) {
                    const uint32_t sp = S.longl[3 * s], dp = S.longl[3 * s + 1], ln = S.longl[3 * s + 2];
                    for (uint32_t i = lane; i < ln; i += 32) dst[dp + i] = S.in[sp + i];
                }
                const uint32_t lastq = S.seqpos[nseq - 1];
                anchor = t0 + lastq + S.len[lastq];
                out_pos += total;
                CTA_SYNC();
                if (tid == 0) S.nlong = 0;
            }
        }

        // ---------------- last literals
        {
            const uint32_t lit = n - anchor;
            const uint32_t fin = out_pos + 1 + lit + (lit >= 15 ? 1 + (lit - 15) / 255 : 0);
            if (fin >= n) { if (tid == 0) blk_csize[blk] = n | 0x80000000u; }   // stored block (LZ4F rule)
            else {
                uint8_t* op = dst + out_pos;
                uint32_t hl = 1;
                if (lit >= 15) hl += 1 + (lit - 15) / 255;
                if (tid == 0) {
                    uint8_t* q = op;
                    *q++ = (uint8_t)((lit >= 15 ? 15u : lit) << 4);
                    if (lit >= 15) { uint32_t x = lit - 15; while (x >= 255) { *q++ = 255; x -= 255; } *q++ = (uint8_t)x; }
                    blk_csize[blk] = fin;
                }
                for (uint32_t i = tid; i < lit; i += C_NT) op[hl + i] = S.in[anchor + i];
            }
        }
    }
}

// ============================================================================ XXH32 (content checksum)
// 4 consecutive lanes own the 4 accumulators of one buffer; 8 buffers per warp.
__global__ void __launch_bounds__(128)
xxh32_kernel(const uint8_t* __restrict__ base, const uint64_t* __restrict__ offs, const uint64_t* __restrict__ lens,
             const uint32_t* __restrict__ lens32, uint64_t stride, uint64_t total_bytes, uint32_t* __restrict__ out, uint32_t nbuf)
{
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, j = threadIdx.x & 3, lane = threadIdx.x & 31;
    const bool live = g < nbuf;
    uint64_t off = 0, n = 0;
    if (live) {
        if (offs) { off = offs[g]; n = lens ? lens[g] : offs[g + 1] - off; }
        else { off = (uint64_t)g * stride; n = zmt_chunk_len(lens32, g, total_bytes, (uint32_t)stride); }
    }
    const uint8_t* p = base + off;
    uint32_t acc = j == 0 ? XXP1 + XXP2 : j == 1 ? XXP2 : j == 2 ? 0u : 0u - XXP1;
    const uint64_t ns = n >> 4;
    if (((uintptr_t)p & 3) == 0) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(p) + j;
        uint64_t s = 0;
        for (; s + 8 <= ns; s += 8) {
            uint32_t x[8];
#pragma unroll
            for (int u = 0; u < 8; u++) x[u] = __ldg(w + 4 * (s + u));
#pragma unroll
            for (int u = 0; u < 8; u++) acc = xxh32_round(acc, x[u]);
        }
        for (; s < ns; s++) acc = xxh32_round(acc, __ldg(w + 4 * s));
    } else {
        for (uint64_t s = 0; s < ns; s++) acc = xxh32_round(acc, ldg_le32(p + 16 * s + 4 * j));
    }
    const uint32_t gl = lane & ~3u;
    const uint32_t a1 = __shfl_sync(ZMT_FULL_MASK, acc, gl), a2 = __shfl_sync(ZMT_FULL_MASK, acc, gl + 1);
    const uint32_t a3 = __shfl_sync(ZMT_FULL_MASK, acc, gl + 2), a4 = __shfl_sync(ZMT_FULL_MASK, acc, gl + 3);
    if (live && j == 0) {
        uint32_t h = n >= 16 ? rotl32(a1, 1) + rotl32(a2, 7) + rotl32(a3, 12) + rotl32(a4, 18) : XXP5;
        h += (uint32_t)n;
        const uint8_t* q = p + (ns << 4);
        const uint8_t* end = p + n;
        while (q + 4 <= end) { h = rotl32(h + ldg_le32(q) * XXP3, 17) * XXP4; q += 4; }
        while (q < end) { h = rotl32(h + (*q) * XXP5, 11) * XXP1; q++; }
        h ^= h >> 15; h *= XXP2; h ^= h >> 13; h *= XXP3; h ^= h >> 16;
        out[g] = h;
    }
}

// ============================================================================ frame pack
// frame bytes = 12 (skippable hdr) + 4 magic + 2 (FLG,BD) + [8 content size] + 1 HC
//               + sum(4 + block bytes) + 4 end mark + 4 content checksum
__global__ void lz4_frame_sizes_kernel(const uint32_t* __restrict__ blk_csize, uint64_t in_bytes, uint32_t chunk_size,
                                       const uint32_t* __restrict__ chunk_bytes, uint32_t bpc, uint32_t nchunks, uint64_t* __restrict__ frame_size)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const uint64_t cbytes = zmt_chunk_len(chunk_bytes, c, in_bytes, chunk_size);
    uint64_t sz = 12 + 4 + 2 + (cbytes ? 8 : 0) + 1 + 4 + 4;
    for (uint32_t b = 0; b < bpc; b++) {
        const uint32_t cs = blk_csize[c * bpc + b];
        if (cs) sz += 4 + (cs & 0x7FFFFFFFu);
    }
    frame_size[c] = sz;
}

// Single-CTA exclusive scan of uint64 sizes -> offsets[0..n] (offsets[n] = total).
__global__ void __launch_bounds__(1024) scan_u64_kernel(const uint64_t* __restrict__ sizes, uint64_t* __restrict__ offsets, uint32_t n)
{
    __shared__ unsigned long long ws[33];
    __shared__ unsigned long long carry;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        unsigned long long v = i < n ? sizes[i] : 0, inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned long long y = __shfl_up_sync(ZMT_FULL_MASK, inc, d); if (lane >= (uint32_t)d) inc += y; }
        if (lane == 31) ws[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned long long x = ws[lane], xi = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { unsigned long long y = __shfl_up_sync(ZMT_FULL_MASK, xi, d); if (lane >= (uint32_t)d) xi += y; }
            ws[lane] = xi - x;
            if (lane == 31) ws[32] = xi;
        }
        __syncthreads();
        const unsigned long long c0 = carry;
        if (i < n) offsets[i] = c0 + ws[wid] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = c0 + ws[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = carry;
}

// One CTA per (chunk, block): copies the block payload to its final place; block 0 also writes
// the headers, the last block the end mark + checksum.
__global__ void __launch_bounds__(256)
lz4_frame_pack_kernel(const uint8_t* __restrict__ in, uint64_t in_bytes, uint32_t chunk_size, const uint32_t* __restrict__ chunk_bytes, uint32_t bpc,
                      const uint8_t* __restrict__ tmp, const uint32_t* __restrict__ blk_csize,
                      const uint32_t* __restrict__ chk, const uint64_t* __restrict__ frame_off,
                      uint8_t* __restrict__ out, uint32_t nblocks)
{
    for (uint32_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const uint32_t c = blk / bpc, b = blk % bpc;
        const uint64_t cbase = (uint64_t)c * chunk_size;
        const uint64_t cbytes = zmt_chunk_len(chunk_bytes, c, in_bytes, chunk_size);
        const uint32_t hdr = 12 + 4 + 2 + (cbytes ? 8 : 0) + 1;
        uint8_t* f = out + frame_off[c];
        const uint32_t cs = blk_csize[blk];
        if (cs == 0 && b != 0) continue;                   // block beyond the end of a short chunk
        uint64_t pos = hdr;                                // payload offset of this block inside the frame
        for (uint32_t k = 0; k < b; k++) { const uint32_t x = blk_csize[c * bpc + k]; if (x) pos += 4 + (x & 0x7FFFFFFFu); }
        const bool last = (b + 1 == bpc) || blk_csize[blk + 1] == 0;
        if (threadIdx.x == 0) {
            if (b == 0) {
                const uint64_t fsz = frame_off[c + 1] - frame_off[c];
                uint8_t h[10];
                stg_le32(f, 0x184D2A50u); stg_le32(f + 4, 4); stg_le32(f + 8, (uint32_t)(fsz - 12));
                stg_le32(f + 12, 0x184D2204u);
                h[0] = cbytes ? 0x6C : 0x64; h[1] = 0x40;
                for (int i = 0; i < 8; i++) h[2 + i] = (uint8_t)(cbytes >> (8 * i));
                const uint32_t hl = cbytes ? 10 : 2;
                for (uint32_t i = 0; i < hl; i++) f[16 + i] = h[i];
                f[16 + hl] = (uint8_t)(xxh32_small(h, hl, 0) >> 8);
            }
            if (cs) stg_le32(f + pos, cs);
            if (last) {
                const uint64_t e = pos + (cs ? 4 + (cs & 0x7FFFFFFFu) : 0);
                stg_le32(f + e, 0); stg_le32(f + e + 4, chk[c]);
            }
        }
        if (cs) {
            const uint8_t* s = (cs & 0x80000000u) ? in + cbase + (uint64_t)b * LZ4_BLK : tmp + (uint64_t)blk * ZMT_LZ4_TMP_STRIDE;
            coop_copy_g2g(f + pos + 4, s, cs & 0x7FFFFFFFu, threadIdx.x, blockDim.x);
        }
    }
}

// ============================================================================ decoder
// One warp per frame; all lanes parse the token stream in lockstep (uniform control flow),
// copies are spread over the 32 lanes.  Handles linked and independent blocks, stored
// blocks, block checksums (skipped), content size + content checksum (checked by the
// follow-up xxh32_kernel + lz4_verify_kernel).
#define D_WARPS 4

__device__ __forceinline__ void warp_copy_lit(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane)
{
    if (n >= 64 && ((((uintptr_t)dst) ^ ((uintptr_t)src)) & 3) == 0) {
        // same 4-byte phase: word copies in the middle
        uint32_t head = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
        if (lane < head) dst[lane] = src[lane];
        const uint32_t nw = (n - head) >> 2;
        const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src + head);
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + head);
        for (uint32_t i = lane; i < nw; i += 32) d4[i] = s4[i];
        for (uint32_t i = head + (nw << 2) + lane; i < n; i += 32) dst[i] = src[i];
    } else {
        for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
    }
}

// returns decoded size or 0xFFFFFFFF on error.  `hist` = bytes of valid history before dst.
__device__ uint32_t warp_decode_block(const uint8_t* __restrict__ src, uint32_t srcSize, uint8_t* dst, uint32_t dstCap,
                                      uint64_t hist, uint32_t lane)
{
    uint32_t ip = 0, op = 0;
    if (srcSize == 0) return 0xFFFFFFFFu;
    for (;;) {
        if (ip >= srcSize) return 0xFFFFFFFFu;
        const uint32_t token = src[ip++];
        uint32_t lit = token >> 4;
        if (lit == 15) {
            uint32_t b;
            do { if (ip >= srcSize) return 0xFFFFFFFFu; b = src[ip++]; lit += b; } while (b == 255);
        }
        if (lit > srcSize - ip || lit > dstCap - op) return 0xFFFFFFFFu;
        warp_copy_lit(dst + op, src + ip, lit, lane);
        ip += lit; op += lit;
        if (ip == srcSize) break;
        if (srcSize - ip < 2) return 0xFFFFFFFFu;
        const uint32_t off = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8);
        ip += 2;
        if (off == 0 || (uint64_t)off > (uint64_t)op + hist) return 0xFFFFFFFFu;
        uint32_t ml = token & 15;
        if (ml == 15) {
            uint32_t b;
            do { if (ip >= srcSize) return 0xFFFFFFFFu; b = src[ip++]; ml += b; } while (b == 255);
        }
        ml += 4;
        if (ml > dstCap - op) return 0xFFFFFFFFu;
        __syncwarp();                                      // literals (and earlier matches) visible to all lanes
        uint8_t* d = dst + op;
        const uint8_t* m = d - off;
        if (off >= ml) { for (uint32_t i = lane; i < ml; i += 32) d[i] = m[i]; }
        else if (off >= 32) {                              // overlapping, period >= warp width: 32-byte waves
            for (uint32_t i = 0; i < ml; i += 32) { if (i + lane < ml) d[i + lane] = m[i + lane]; __syncwarp(); }
        } else {                                           // short period: replicate the pattern
            for (uint32_t i = lane; i < ml; i += 32) d[i] = m[i % off];
        }
        op += ml;
        __syncwarp();
    }
    return op;
}

__global__ void __launch_bounds__(32 * D_WARPS)
lz4_decode_frames_kernel(const uint8_t* __restrict__ in, const uint64_t* __restrict__ frame_off, const uint32_t* __restrict__ frame_csize,
                         uint8_t* __restrict__ out, const uint64_t* __restrict__ out_off, uint64_t* __restrict__ out_size,
                         uint32_t* __restrict__ status, uint32_t* __restrict__ stored_chk, uint32_t nframes)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t f = blockIdx.x * D_WARPS + (threadIdx.x >> 5);
    if (f >= nframes) return;
    const uint8_t* p = in + frame_off[f] + 12;              // LZ4F frame (after the skippable header)
    const uint32_t fs = frame_csize[f];
    uint8_t* dst = out + out_off[f];
    const uint64_t cap = out_off[f + 1] - out_off[f];
    uint32_t st = ZMT_ST_OK, has_chk = 0, chkv = 0;
    uint64_t total = 0;
    do {
        if (fs < 7 + 4) { st = ZMT_ST_TRUNCATED; break; }
        if (ldg_le32(p) != 0x184D2204u) { st = ZMT_ST_BAD_MAGIC; break; }
        const uint32_t flg = p[4], bd = p[5];
        if ((flg >> 6) != 1 || (flg & 2) || (bd & 0x8F) || ((bd >> 4) & 7) < 4) { st = ZMT_ST_BAD_HEADER; break; }
        const uint32_t indep = (flg >> 5) & 1, bchk = (flg >> 4) & 1, csz = (flg >> 3) & 1, cchk = (flg >> 2) & 1, did = flg & 1;
        const uint32_t blkmax = 1u << (8 + 2 * ((bd >> 4) & 7));
        const uint32_t hl = 2 + (csz ? 8 : 0) + (did ? 4 : 0);
        if (fs < 4 + hl + 1 + 4) { st = ZMT_ST_TRUNCATED; break; }
        {
            uint8_t h[14];
            for (uint32_t i = 0; i < hl; i++) h[i] = p[4 + i];
            if (((xxh32_small(h, hl, 0) >> 8) & 0xFF) != p[4 + hl]) { st = ZMT_ST_HDR_CHECKSUM; break; }
        }
        const uint64_t content = csz ? ldg_le64(p + 6) : 0;
        uint32_t ip = 4 + hl + 1;
        for (;;) {
            if (fs - ip < 4) { st = ZMT_ST_TRUNCATED; break; }
            const uint32_t bh = ldg_le32(p + ip); ip += 4;
            if (bh == 0) break;
            const uint32_t bs = bh & 0x7FFFFFFFu;
            if (bs > blkmax) { st = ZMT_ST_BLOCK; break; }
            if (fs - ip < bs + (bchk ? 4 : 0)) { st = ZMT_ST_TRUNCATED; break; }
            if (bh & 0x80000000u) {
                if (bs > cap - total) { st = ZMT_ST_DST_SMALL; break; }
                __syncwarp();
                warp_copy_lit(dst + total, p + ip, bs, lane);
                total += bs;
            } else {
                const uint64_t room = cap - total;
                const uint32_t dcap = room < blkmax ? (uint32_t)room : blkmax;
                const uint64_t hist = indep ? 0 : (total < 65536 ? total : 65536);
                const uint32_t d = warp_decode_block(p + ip, bs, dst + total, dcap, hist, lane);
                if (d == 0xFFFFFFFFu) { st = ZMT_ST_BLOCK; break; }
                total += d;
            }
            ip += bs + (bchk ? 4 : 0);
            __syncwarp();
        }
        if (st != ZMT_ST_OK) break;
        if (cchk) {
            if (fs - ip < 4) { st = ZMT_ST_TRUNCATED; break; }
            has_chk = 1; chkv = ldg_le32(p + ip); ip += 4;
        }
        if (csz && content != total) { st = ZMT_ST_CONTENT_SIZE; break; }
        if (ip != fs) { st = ZMT_ST_TRAILING; break; }
    } while (0);
    if (lane == 0) { status[f] = st | (has_chk ? ZMT_ST_HAS_CHK : 0); stored_chk[f] = chkv; out_size[f] = total; }
}

// compares xxh32 of the decoded output with the stored content checksum
__global__ void lz4_verify_kernel(uint32_t* __restrict__ status, const uint32_t* __restrict__ stored_chk,
                                  const uint32_t* __restrict__ computed, uint32_t nframes)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    uint32_t st = status[f];
    if ((st & ZMT_ST_HAS_CHK) && (st & 0xFF) == ZMT_ST_OK && stored_chk[f] != computed[f]) st = (st & ~0xFFu) | ZMT_ST_CONTENT_CHECKSUM;
    status[f] = st & 0xFF;
}

// ============================================================================ host launchers
// ZSTDMT_B200_DEBUG_SYNC=1: synchronise after every kernel and name the one that failed (debug only)
static bool zmt_dbg_check(cudaStream_t st, const char* what)
{
    static int on = -1;
    if (on < 0) on = getenv("ZSTDMT_B200_DEBUG_SYNC") ? 1 : 0;
    if (!on) return true;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "[zstdmt_b200] %s failed: %s\n", what, cudaGetErrorString(e)); return false; }
    return true;
}

static inline int zmt_sm_count()
{
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

extern "C" uint32_t zmt_chunk_count(uint64_t in_bytes, uint32_t chunk_size)
{
    if (chunk_size == 0) return 0;
    const uint64_t n = in_bytes ? (in_bytes + chunk_size - 1) / chunk_size : 1;   // empty input still yields one frame
    return n > 0x7FFFFFFFull ? 0 : (uint32_t)n;
}

extern "C" size_t zmt_lz4c_workspace_bytes(uint32_t nchunks, uint32_t chunk_size)
{
    if (chunk_size == 0) return 0;
    const uint64_t bpc = ((uint64_t)chunk_size + LZ4_BLK - 1) / LZ4_BLK;
    const uint64_t nblocks = (uint64_t)nchunks * bpc;
    uint64_t sz = 0;
    sz += nblocks * ZMT_LZ4_TMP_STRIDE;                  // per-block temp slots
    sz += ((nblocks * 4 + 255) & ~255ull);               // blk_csize
    sz += (((uint64_t)nchunks * 4 + 255) & ~255ull);     // chk
    sz += ((((uint64_t)nchunks + 1) * 8 + 255) & ~255ull); // frame_size
    return (size_t)sz + 1024;
}

extern "C" uint64_t zmt_lz4c_out_bound(uint32_t nchunks, uint32_t chunk_size)
{
    if (chunk_size == 0) return 0;
    const uint64_t bpc = ((uint64_t)chunk_size + LZ4_BLK - 1) / LZ4_BLK;
    return (uint64_t)nchunks * ((uint64_t)chunk_size + 12 + 15 + 8 + 4 * bpc) + 256;
}

extern "C" int zmt_lz4_compress_device(const void* d_in, uint64_t in_bytes, uint32_t chunk_size, const uint32_t* d_chunk_bytes,
                                       uint32_t nchunks, void* d_work, void* d_out, uint64_t* d_frame_off, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (chunk_size == 0 || nchunks == 0) return ZMT_ST_BAD_ARG;
    if (!d_chunk_bytes && nchunks != zmt_chunk_count(in_bytes, chunk_size)) return ZMT_ST_BAD_ARG;
    const uint32_t bpc = (uint32_t)(((uint64_t)chunk_size + LZ4_BLK - 1) / LZ4_BLK);
    if ((uint64_t)nchunks * bpc > 0x7FFFFFFFull) return ZMT_ST_BAD_ARG;
    const uint32_t nblocks = nchunks * bpc;
    uint8_t* w = (uint8_t*)d_work;
    uint8_t* tmp = w; w += (uint64_t)nblocks * ZMT_LZ4_TMP_STRIDE;
    uint32_t* blk_csize = (uint32_t*)w; w += (((uint64_t)nblocks * 4 + 255) & ~255ull);
    uint32_t* chk = (uint32_t*)w; w += (((uint64_t)nchunks * 4 + 255) & ~255ull);
    uint64_t* frame_size = (uint64_t*)w;

    cudaFuncSetAttribute(lz4_compress_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CompressSmem));
    const unsigned dm = getenv("ZSTDMT_B200_DEBUG_MASK") ? (unsigned)atoi(getenv("ZSTDMT_B200_DEBUG_MASK")) : 31u;
    const uint32_t maxc = (uint32_t)(zmt_sm_count() * 2 * 8);
    uint32_t gridc = nblocks < maxc ? nblocks : maxc;
    if (getenv("ZSTDMT_B200_DEBUG_GRID")) gridc = (uint32_t)atoi(getenv("ZSTDMT_B200_DEBUG_GRID"));
    if (dm & 1) lz4_compress_blocks_kernel<<<gridc, C_NT, sizeof(CompressSmem), stream>>>((const uint8_t*)d_in, in_bytes, chunk_size, d_chunk_bytes, bpc, tmp, blk_csize, nblocks, (getenv("ZSTDMT_B200_NO_TMA") ? 1u : 0u) | (getenv("ZSTDMT_B200_DEBUG_FLAGS") ? (unsigned)atoi(getenv("ZSTDMT_B200_DEBUG_FLAGS")) : 0u));
    zmt_dbg_check(stream, "lz4_compress_blocks_kernel");
    if (dm & 2) xxh32_kernel<<<(nchunks * 4 + 127) / 128, 128, 0, stream>>>((const uint8_t*)d_in, nullptr, nullptr, d_chunk_bytes, chunk_size, in_bytes, chk, nchunks);
    zmt_dbg_check(stream, "xxh32_kernel");
    if (dm & 4) lz4_frame_sizes_kernel<<<(nchunks + 255) / 256, 256, 0, stream>>>(blk_csize, in_bytes, chunk_size, d_chunk_bytes, bpc, nchunks, frame_size);
    zmt_dbg_check(stream, "lz4_frame_sizes_kernel");
    if (dm & 8) scan_u64_kernel<<<1, 1024, 0, stream>>>(frame_size, d_frame_off, nchunks);
    zmt_dbg_check(stream, "scan_u64_kernel");
    const uint32_t maxp = (uint32_t)(zmt_sm_count() * 16);
    const uint32_t gridp = nblocks < maxp ? nblocks : maxp;
    if (dm & 16) lz4_frame_pack_kernel<<<gridp, 256, 0, stream>>>((const uint8_t*)d_in, in_bytes, chunk_size, d_chunk_bytes, bpc, tmp, blk_csize, chk, d_frame_off, (uint8_t*)d_out, nblocks);
    zmt_dbg_check(stream, "lz4_frame_pack_kernel");
    return cudaGetLastError() == cudaSuccess ? ZMT_ST_OK : ZMT_ST_CUDA;
}

extern "C" size_t zmt_lz4d_workspace_bytes(uint32_t nframes)
{
    return (size_t)((((uint64_t)nframes * 4 + 255) & ~255ull) * 2 + 1024);
}

extern "C" int zmt_lz4_decompress_device(const void* d_in, const uint64_t* d_frame_off, const uint32_t* d_frame_csize, uint32_t nframes,
                                         void* d_out, const uint64_t* d_out_off, uint64_t* d_out_size, uint32_t* d_status,
                                         void* d_work, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nframes == 0) return ZMT_ST_OK;
    uint8_t* w = (uint8_t*)d_work;
    uint32_t* stored = (uint32_t*)w; w += (((uint64_t)nframes * 4 + 255) & ~255ull);
    uint32_t* computed = (uint32_t*)w;
    lz4_decode_frames_kernel<<<(nframes + D_WARPS - 1) / D_WARPS, 32 * D_WARPS, 0, stream>>>((const uint8_t*)d_in, d_frame_off, d_frame_csize,
                                                                                   (uint8_t*)d_out, d_out_off, d_out_size, d_status, stored, nframes);
    xxh32_kernel<<<(nframes * 4 + 127) / 128, 128, 0, stream>>>((const uint8_t*)d_out, d_out_off, d_out_size, nullptr, 0, 0, computed, nframes);
    lz4_verify_kernel<<<(nframes + 255) / 256, 256, 0, stream>>>(d_status, stored, computed, nframes);
    return cudaGetLastError() == cudaSuccess ? ZMT_ST_OK : ZMT_ST_CUDA;
}
