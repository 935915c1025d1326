// zstd_decode.cu — Zstandard frame decoding on sm_100a (the ZSTD_decompressStream step of
// /root/reference/lib/zstd-mt_decompress.c:442-527, one frame per 12-byte container header).
//
// Three kernels per batch, all block-parallel (a zstd frame of a 1 MiB chunk has ~64 blocks):
//   zstd_entropy_kernel   one warp per block: literals (raw / RLE / Huffman with direct weights, 1 or 4 streams —
//                         lanes 0..3 own the streams) and sequences (predefined FSE tables, lane 4) are decoded
//                         into per-block scratch; the block's regenerated size = literals + sum of match lengths
//   zstd_offsets_kernel   one thread per frame: exclusive scan of the regenerated sizes -> output offset per block,
//                         content-size check against the frame header
//   zstd_execute_kernel   one warp per block: literal / match copies in sequence order; a match that reaches below
//                         the block's own output waits for the `done` flags of the blocks it reads from
// Host side (zmt_zstd_scan_host): walks frame + block headers (3 bytes per block) and sizes the scratch.
//
// Scope: everything our encoder emits plus raw / RLE blocks and 1-stream literals.  Streams that need
// FSE-described tables, FSE-coded Huffman weights, treeless literals or repeat offsets (what libzstd emits for the
// reference) are reported per frame as ZMT_ST_UNSUPPORTED — never decoded on the CPU.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "zmt_dev.h"

// ---------------------------------------------------------------- block descriptors (host-built)
#define ZB_RAW 0u
#define ZB_RLE 1u
#define ZB_CMP 2u
struct ZBlk {
    uint64_t comp_off;      // offset of the block CONTENT (after its 3-byte header) in the input buffer
    uint32_t comp_size;     // content bytes (1 for RLE)
    uint32_t frame;         // frame index in the batch
    uint32_t type;          // ZB_*
    uint32_t regen_hint;    // raw / RLE: regenerated size; compressed: literal bytes (regenerated)
    uint32_t nseq;          // compressed: number of sequences
    uint32_t first;         // 1 if first block of its frame
    uint64_t seq_off;       // scratch offsets (bytes) for this block's sequence records / literals
    uint64_t lit_off;
};
struct ZDSeq { uint32_t ll; uint32_t off; uint32_t ml; uint32_t pad; };

// ---------------------------------------------------------------- predefined FSE decode tables
struct ZFseDTable { uint8_t sym[64]; uint8_t nb[64]; uint16_t base[64]; uint32_t log; };
__constant__ ZFseDTable d_fse_ll, d_fse_of, d_fse_ml;
__constant__ uint32_t d_ll_base[36], d_ml_base[53];
__constant__ uint8_t  d_ll_bits[36], d_ml_bits[53];

// ---------------------------------------------------------------- backward bit reader over global memory
// bit `off` = number of unread bits; reads below the stream start return zeros (off goes negative = exhausted)
struct BackBits {
    const uint8_t* p; int32_t off;
    __device__ __forceinline__ bool init(const uint8_t* s, uint32_t n)
    {
        if (n == 0) return false;
        const uint32_t lastb = s[n - 1];
        if (lastb == 0) return false;
        p = s; off = (int32_t)(n * 8) - (int32_t)(__clz(lastb) - 24 + 1);
        return true;
    }
    // up to 25 bits
    __device__ __forceinline__ uint32_t read(uint32_t nb)
    {
        if (nb == 0) return 0;
        off -= (int32_t)nb;
        const int32_t start = off;                       // may be negative
        const int32_t byte0 = start >= 0 ? (start >> 3) : -((-start + 7) >> 3);
        const int32_t sh = start - byte0 * 8;            // 0..7
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { const int32_t b = byte0 + k; if (b >= 0) v |= (uint32_t)p[b] << (8 * k); }
        // bits above the stream end never matter: callers only ask for bits that exist or pad below the start
        return (v >> sh) & ((1u << nb) - 1);
    }
};

__device__ __forceinline__ void zd_fail(uint32_t* status, uint32_t f, uint32_t code) { atomicCAS(&status[f], 0u, code); }

// ---------------------------------------------------------------- kernel 1: entropy decode
#define ZD_WARPS 4
__global__ void __launch_bounds__(32 * ZD_WARPS)
zstd_entropy_kernel(const uint8_t* __restrict__ in, const ZBlk* __restrict__ blocks, uint32_t nblocks,
                    uint8_t* __restrict__ scratch, uint32_t* __restrict__ regen, uint32_t* __restrict__ status)
{
    __shared__ uint16_t htab[ZD_WARPS][2048];            // (nbBits << 8) | symbol
    __shared__ uint8_t  wts[ZD_WARPS][256];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x * ZD_WARPS + wid;
    if (b >= nblocks) return;
    const ZBlk B = blocks[b];
    if (B.type != ZB_CMP) { if (lane == 0) regen[b] = B.regen_hint; return; }
    const uint8_t* src = in + B.comp_off;
    const uint32_t n = B.comp_size;
    uint8_t* lit = scratch + B.lit_off;
    ZDSeq* seqs = reinterpret_cast<ZDSeq*>(scratch + B.seq_off);
    uint16_t* T = htab[wid];

    // ---- literals section
    if (n < 1) { if (lane == 0) zd_fail(status, B.frame, ZMT_ST_BLOCK); return; }
    const uint32_t b0 = src[0], ltype = b0 & 3, sf = (b0 >> 2) & 3;
    uint32_t lregen, lcomp = 0, lhdr, streams = 1;
    if (ltype < 2) {
        if (sf == 0 || sf == 2) { lregen = b0 >> 3; lhdr = 1; }
        else if (sf == 1) { lregen = (b0 >> 4) | ((uint32_t)src[1] << 4); lhdr = 2; }
        else { lregen = (b0 >> 4) | ((uint32_t)src[1] << 4) | ((uint32_t)src[2] << 12); lhdr = 3; }
        lcomp = ltype == 0 ? lregen : 1;
    } else {
        if (ltype == 3) { if (lane == 0) zd_fail(status, B.frame, ZMT_ST_UNSUPPORTED); return; }    // treeless literals
        if (sf < 2) { const uint32_t v = src[0] | (src[1] << 8) | ((uint32_t)src[2] << 16); lregen = (v >> 4) & 0x3FF; lcomp = (v >> 14) & 0x3FF; lhdr = 3; streams = sf == 0 ? 1 : 4; }
        else if (sf == 2) { const uint32_t v = ldg_le32(src); lregen = (v >> 4) & 0x3FFF; lcomp = (v >> 18) & 0x3FFF; lhdr = 4; streams = 4; }
        else { const uint64_t v = (uint64_t)ldg_le32(src) | ((uint64_t)src[4] << 32); lregen = (uint32_t)((v >> 4) & 0x3FFFF); lcomp = (uint32_t)((v >> 22) & 0x3FFFF); lhdr = 5; streams = 4; }
    }
    if (lhdr + lcomp > n || lregen != B.regen_hint) { if (lane == 0) zd_fail(status, B.frame, ZMT_ST_BLOCK); return; }
    const uint8_t* lp = src + lhdr;
    if (ltype == 0) { for (uint32_t i = lane; i < lregen; i += 32) lit[i] = lp[i]; }
    else if (ltype == 1) { const uint8_t v = lp[0]; for (uint32_t i = lane; i < lregen; i += 32) lit[i] = v; }
    else {
        // Huffman tree: direct 4-bit weights only
        const uint32_t hb = lp[0];
        if (hb < 128) { if (lane == 0) zd_fail(status, B.frame, ZMT_ST_UNSUPPORTED); return; }     // FSE-coded weights
        const uint32_t nw = hb - 127, tbytes = 1 + (nw + 1) / 2;
        if (tbytes + (streams == 4 ? 6u : 0u) > lcomp) { if (lane == 0) zd_fail(status, B.frame, ZMT_ST_BLOCK); return; }
        uint32_t total = 0;
        for (uint32_t i = lane; i < 256; i += 32) {
            uint32_t w = 0;
            if (i < nw) { const uint32_t by = lp[1 + i / 2]; w = (i & 1) ? (by & 15) : (by >> 4); }
            wts[wid][i] = (uint8_t)w;
            if (w) total += 1u << (w - 1);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) total += __shfl_xor_sync(ZMT_FULL_MASK, total, d);
        if (total == 0 || total >= 2048) { if (lane == 0) zd_fail(status, B.frame, ZMT_ST_BLOCK); return; }
        const uint32_t maxbits = 32 - __clz(total);       // highbit(total) + 1
        const uint32_t left = (1u << maxbits) - total;
        if (left == 0 || (left & (left - 1)) || maxbits > 11) { if (lane == 0) zd_fail(status, B.frame, ZMT_ST_BLOCK); return; }
        __syncwarp();
        if (lane == 0) wts[wid][nw] = (uint8_t)(32 - __clz(left));          // implied last weight = highbit(left) + 1
        __syncwarp();
        // table: ascending weight, then symbol order (RFC 8878 §4.2.1.3); lane owns symbols lane, lane+32, ...
        // start index of symbol s = sum over (w' < w) cnt[w'] << (w'-1)  +  rank among equal weights << (w-1)
        uint32_t cntw[12];
#pragma unroll
        for (int w = 0; w < 12; w++) cntw[w] = 0;
        for (uint32_t s = 0; s <= nw; s++) { const uint32_t w = wts[wid][s]; if (w < 12) cntw[w]++; else cntw[0]++; }   // uniform loop, every lane
        for (uint32_t s = lane; s <= nw; s += 32) {
            const uint32_t w = wts[wid][s];
            if (!w || w > 11) continue;
            uint32_t start = 0;
            for (uint32_t ww = 1; ww < w; ww++) start += cntw[ww] << (ww - 1);
            uint32_t r = 0;
            for (uint32_t t = 0; t < s; t++) r += wts[wid][t] == w ? 1u : 0u;
            start += r << (w - 1);
            const uint16_t e = (uint16_t)(((maxbits + 1 - w) << 8) | s);
            for (uint32_t k = 0; k < (1u << (w - 1)); k++) T[start + k] = e;
        }
        __syncwarp();
        const uint8_t* sp = lp + tbytes;
        uint32_t ssz[4], spos[4], per = lregen, nsym[4];
        if (streams == 4) {
            const uint32_t s1 = sp[0] | (sp[1] << 8), s2 = sp[2] | (sp[3] << 8), s3 = sp[4] | (sp[5] << 8);
            const uint32_t body = lcomp - tbytes - 6;
            if (s1 + s2 + s3 > body) { if (lane == 0) zd_fail(status, B.frame, ZMT_ST_BLOCK); return; }
            ssz[0] = s1; ssz[1] = s2; ssz[2] = s3; ssz[3] = body - s1 - s2 - s3;
            spos[0] = 0; spos[1] = s1; spos[2] = s1 + s2; spos[3] = s1 + s2 + s3;
            per = (lregen + 3) / 4;
            if (per * 3 > lregen) { if (lane == 0) zd_fail(status, B.frame, ZMT_ST_BLOCK); return; }
            nsym[0] = nsym[1] = nsym[2] = per; nsym[3] = lregen - 3 * per;
            sp += 6;
        } else { ssz[0] = lcomp - tbytes; spos[0] = 0; nsym[0] = lregen; ssz[1] = ssz[2] = ssz[3] = 0; spos[1] = spos[2] = spos[3] = 0; nsym[1] = nsym[2] = nsym[3] = 0; }
        bool okh = true;
        if (lane < streams) {
            BackBits R;
            const uint32_t cnt = nsym[lane];
            uint8_t* o = lit + lane * per;
            if (!R.init(sp + spos[lane], ssz[lane])) okh = (cnt == 0 && ssz[lane] == 0) ? false : false;
            else {
                uint32_t st = R.read(maxbits);
                for (uint32_t i = 0; i < cnt; i++) {
                    const uint32_t e = T[st], nb = e >> 8;
                    o[i] = (uint8_t)e;
                    st = ((st << nb) & ((1u << maxbits) - 1)) | R.read(nb);
                }
                if (R.off != -(int32_t)maxbits) okh = false;
            }
        }
        if (!__all_sync(ZMT_FULL_MASK, okh)) { if (lane == 0) zd_fail(status, B.frame, ZMT_ST_BLOCK); return; }
    }

    // ---- sequences section (lane 0 walks the three interleaved FSE states; predefined tables only)
    const uint8_t* qp = src + lhdr + lcomp;
    uint32_t qn = n - lhdr - lcomp;
    uint32_t okq = 1, total_ml = 0;
    if (lane == 0) {
        do {
            if (qn < 1) { okq = 0; break; }
            uint32_t nseq; const uint32_t q0 = qp[0];
            uint32_t used = 1;
            if (q0 == 0) nseq = 0;
            else if (q0 < 128) nseq = q0;
            else if (q0 < 255) { if (qn < 2) { okq = 0; break; } nseq = ((q0 - 128) << 8) + qp[1]; used = 2; }
            else { if (qn < 3) { okq = 0; break; } nseq = qp[1] + (qp[2] << 8) + 0x7F00; used = 3; }
            if (nseq != B.nseq) { okq = 0; break; }
            if (nseq == 0) { if (used != qn) okq = 0; break; }
            if (qn < used + 1) { okq = 0; break; }
            if (qp[used] != 0) { okq = 2; break; }                               // non-predefined table modes
            used++;
            BackBits R;
            if (!R.init(qp + used, qn - used)) { okq = 0; break; }
            uint32_t sLL = R.read(d_fse_ll.log), sOF = R.read(d_fse_of.log), sML = R.read(d_fse_ml.log);
            for (uint32_t i = 0; i < nseq; i++) {
                const uint32_t ofc = d_fse_of.sym[sOF], mlc = d_fse_ml.sym[sML], llc = d_fse_ll.sym[sLL];
                if (ofc > 24) { okq = 0; break; }
                const uint32_t ofv = (1u << ofc) + R.read(ofc);
                const uint32_t ml = d_ml_base[mlc] + R.read(d_ml_bits[mlc]);
                const uint32_t ll = d_ll_base[llc] + R.read(d_ll_bits[llc]);
                if (ofv <= 3) { okq = 2; break; }                                 // repeat offsets: not in the B200 subset
                if (i + 1 < nseq) {
                    sLL = d_fse_ll.base[sLL] + R.read(d_fse_ll.nb[sLL]);
                    sML = d_fse_ml.base[sML] + R.read(d_fse_ml.nb[sML]);
                    sOF = d_fse_of.base[sOF] + R.read(d_fse_of.nb[sOF]);
                }
                if (R.off < 0) { okq = 0; break; }
                ZDSeq q; q.ll = ll; q.off = ofv - 3; q.ml = ml; q.pad = 0;
                seqs[i] = q;
                total_ml += ml;
            }
            if (okq == 1 && R.off != 0) okq = 0;
        } while (0);
        if (okq != 1) zd_fail(status, B.frame, okq == 2 ? ZMT_ST_UNSUPPORTED : ZMT_ST_BLOCK);
        else regen[b] = lregen + total_ml;
    }
}

// ---------------------------------------------------------------- kernel 2: output offsets per block
__global__ void zstd_offsets_kernel(const ZBlk* __restrict__ blocks, uint32_t nblocks, const uint32_t* __restrict__ frame_first_blk,
                                    const uint32_t* __restrict__ regen, uint64_t* __restrict__ blk_out, const uint64_t* __restrict__ out_off,
                                    const uint64_t* __restrict__ expect, unsigned long long* __restrict__ out_size, uint32_t* __restrict__ status,
                                    uint32_t nframes)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const uint32_t b0 = frame_first_blk[f], b1 = frame_first_blk[f + 1];
    uint64_t pos = 0;
    const uint64_t cap = out_off[f + 1] - out_off[f];
    for (uint32_t b = b0; b < b1; b++) { blk_out[b] = out_off[f] + pos; pos += regen[b]; }
    out_size[f] = pos;
    if ((status[f] & 0xFF) == 0 && (pos != expect[f] || pos > cap)) status[f] = pos > cap ? ZMT_ST_DST_SMALL : ZMT_ST_CONTENT_SIZE;
}

// ---------------------------------------------------------------- kernel 3: sequence execution
#define ZX_WARPS 8
__global__ void __launch_bounds__(32 * ZX_WARPS)
zstd_execute_kernel(const uint8_t* __restrict__ in, const ZBlk* __restrict__ blocks, uint32_t nblocks, const uint8_t* __restrict__ scratch,
                    const uint32_t* __restrict__ regen, const uint64_t* __restrict__ blk_out, const uint64_t* __restrict__ out_off,
                    uint8_t* __restrict__ out, uint32_t* __restrict__ done, uint32_t* __restrict__ status)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * ZX_WARPS + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const ZBlk B = blocks[b];
    volatile uint32_t* vdone = done;
    // a frame that already failed: do not touch memory, just release the waiters
    if ((status[B.frame] & 0xFF) != 0) { __syncwarp(); if (lane == 0) { __threadfence(); vdone[b] = 1; } return; }
    uint8_t* dst = out + blk_out[b];
    const uint64_t frame_base = out_off[B.frame];
    const uint32_t rg = regen[b];
    if (B.type == ZB_RAW) { const uint8_t* s = in + B.comp_off; for (uint32_t i = lane; i < rg; i += 32) dst[i] = s[i]; }
    else if (B.type == ZB_RLE) { const uint8_t v = in[B.comp_off]; for (uint32_t i = lane; i < rg; i += 32) dst[i] = v; }
    else {
        const uint8_t* lit = scratch + B.lit_off;
        const ZDSeq* seqs = reinterpret_cast<const ZDSeq*>(scratch + B.seq_off);
        uint32_t op = 0, lp = 0;
        bool bad = false;
        uint32_t waited_to = b;                              // blocks [waited_to, b) are known complete
        for (uint32_t i = 0; i < B.nseq; i++) {
            const ZDSeq q = seqs[i];
            if ((uint64_t)op + q.ll + q.ml > rg || lp + q.ll > B.regen_hint) { bad = true; break; }
            for (uint32_t k = lane; k < q.ll; k += 32) dst[op + k] = lit[lp + k];
            op += q.ll; lp += q.ll;
            const uint64_t abs_pos = (blk_out[b] - frame_base) + op;
            if (q.off == 0 || q.off > abs_pos) { bad = true; break; }
            if (q.off > op) {
                // the match starts below this block: every earlier block it touches must be finished
                const uint64_t need = blk_out[b] + op - q.off;                   // absolute output address of the first source byte
                while (waited_to > 0 && blocks[waited_to - 1].frame == B.frame && blk_out[waited_to - 1] + regen[waited_to - 1] > need) {
                    waited_to--;
                    if (lane == 0) { while (vdone[waited_to] == 0) __nanosleep(64); }
                }
                __syncwarp();
                __threadfence();
            }
            __syncwarp();
            uint8_t* d = dst + op;
            const uint8_t* m = d - q.off;
            const uint32_t ml = q.ml;
            if (q.off >= ml) { for (uint32_t k = lane; k < ml; k += 32) d[k] = m[k]; }
            else if (q.off >= 32) { for (uint32_t k = 0; k < ml; k += 32) { if (k + lane < ml) d[k + lane] = m[k + lane]; __syncwarp(); } }
            else { for (uint32_t k = lane; k < ml; k += 32) d[k] = m[k % q.off]; }
            op += ml;
            __syncwarp();
        }
        if (!bad) {
            const uint32_t rest = B.regen_hint - lp;
            if (op + rest != rg) bad = true;
            else for (uint32_t k = lane; k < rest; k += 32) dst[op + k] = lit[lp + k];
        }
        if (bad && lane == 0) zd_fail(status, B.frame, ZMT_ST_BLOCK);
    }
    __syncwarp();
    if (lane == 0) { __threadfence(); vdone[b] = 1; }
}

// ================================================================ host side
static void zd_build_dtable(ZFseDTable& T, const int16_t* norm, int nsym, int log)
{
    const int size = 1 << log, step = (size >> 1) + (size >> 3) + 3;
    int high = size - 1, pos = 0; uint16_t next[64];
    memset(&T, 0, sizeof(T)); T.log = (uint32_t)log;
    for (int s = 0; s < nsym; s++) { if (norm[s] == -1) { T.sym[high--] = (uint8_t)s; next[s] = 1; } else next[s] = (uint16_t)norm[s]; }
    for (int s = 0; s < nsym; s++)
        for (int i = 0; i < norm[s]; i++) { T.sym[pos] = (uint8_t)s; do { pos = (pos + step) & (size - 1); } while (pos > high); }
    for (int i = 0; i < size; i++) {
        const uint16_t x = next[T.sym[i]]++;
        int hb = 0; while ((1 << (hb + 1)) <= x) hb++;
        T.nb[i] = (uint8_t)(log - hb);
        T.base[i] = (uint16_t)(((uint32_t)x << T.nb[i]) - size);
    }
}

static int zd_tables_init()
{
    static std::vector<int> done;
    int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess) return ZMT_ST_CUDA;
    for (int d : done) if (d == dev) return ZMT_ST_OK;
    static const int16_t LLn[36] = { 4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1 };
    static const int16_t MLn[53] = { 1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1 };
    static const int16_t OFn[29] = { 1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1 };
    static const uint32_t LLb[36] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,128,256,512,1024,2048,4096,8192,16384,32768,65536 };
    static const uint8_t  LLx[36] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16 };
    static const uint32_t MLb[53] = { 3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,33,34,35,37,39,41,43,47,51,59,67,83,99,131,259,515,1027,2051,4099,8195,16387,32771,65539 };
    static const uint8_t  MLx[53] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16 };
    ZFseDTable t;
    zd_build_dtable(t, LLn, 36, 6); if (cudaMemcpyToSymbol(d_fse_ll, &t, sizeof(t)) != cudaSuccess) return ZMT_ST_CUDA;
    zd_build_dtable(t, OFn, 29, 5); if (cudaMemcpyToSymbol(d_fse_of, &t, sizeof(t)) != cudaSuccess) return ZMT_ST_CUDA;
    zd_build_dtable(t, MLn, 53, 6); if (cudaMemcpyToSymbol(d_fse_ml, &t, sizeof(t)) != cudaSuccess) return ZMT_ST_CUDA;
    cudaMemcpyToSymbol(d_ll_base, LLb, sizeof(LLb)); cudaMemcpyToSymbol(d_ml_base, MLb, sizeof(MLb));
    cudaMemcpyToSymbol(d_ll_bits, LLx, sizeof(LLx)); cudaMemcpyToSymbol(d_ml_bits, MLx, sizeof(MLx));
    if (cudaGetLastError() != cudaSuccess) return ZMT_ST_CUDA;
    done.push_back(dev);
    return ZMT_ST_OK;
}

static inline uint32_t h_rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// Walk one zstd frame on the host (frame header + 3-byte block headers + the two section headers of every
// compressed block).  `base_off` = offset of the frame's first byte inside the batch input buffer.
// Appends block descriptors (scratch offsets assigned from *scratch_used).  Returns ZMT_ST_*.
extern "C" int zmt_zstd_scan_frame_host(const uint8_t* frame, size_t n, uint64_t base_off, uint32_t frame_idx,
                                        void* blocks_out, uint32_t* nblocks_io, uint32_t max_blocks,
                                        uint64_t* scratch_used, uint64_t* content_size)
{
    ZBlk* out = (ZBlk*)blocks_out;
    if (n < 6 || h_rd32(frame) != 0xFD2FB528u) return ZMT_ST_BAD_MAGIC;
    const uint32_t fhd = frame[4], fcs = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
    if (fhd & 0x08) return ZMT_ST_BAD_HEADER;
    if (fhd & 0x04) return ZMT_ST_UNSUPPORTED;                       // content checksum (XXH64) not produced by the reference path
    size_t pos = 5 + (single ? 0 : 1) + (did == 0 ? 0 : did == 1 ? 1 : did == 2 ? 2 : 4);
    const size_t fl = fcs == 0 ? (single ? 1 : 0) : fcs == 1 ? 2 : fcs == 2 ? 4 : 8;
    if (fl == 0) return ZMT_ST_UNSUPPORTED;                           // unknown content size: the reference path always has it
    if (n < pos + fl) return ZMT_ST_TRUNCATED;
    uint64_t cs;
    if (fl == 1) cs = frame[pos]; else if (fl == 2) cs = (uint64_t)(frame[pos] | (frame[pos + 1] << 8)) + 256;
    else if (fl == 4) cs = h_rd32(frame + pos); else cs = (uint64_t)h_rd32(frame + pos) | ((uint64_t)h_rd32(frame + pos + 4) << 32);
    pos += fl;
    *content_size = cs;
    bool first = true;
    for (;;) {
        if (n - pos < 3) return ZMT_ST_TRUNCATED;
        const uint32_t bh = frame[pos] | (frame[pos + 1] << 8) | ((uint32_t)frame[pos + 2] << 16); pos += 3;
        const uint32_t last = bh & 1, type = (bh >> 1) & 3, bs = bh >> 3;
        if (type == 3 || bs > 128 * 1024) return ZMT_ST_BLOCK;
        if (*nblocks_io >= max_blocks) return ZMT_ST_DST_SMALL;
        ZBlk& B = out[*nblocks_io];
        memset(&B, 0, sizeof(B));
        B.frame = frame_idx; B.type = type; B.first = first ? 1 : 0; B.comp_off = base_off + pos;
        if (type == ZB_RAW) { if (n - pos < bs) return ZMT_ST_TRUNCATED; B.comp_size = bs; B.regen_hint = bs; pos += bs; }
        else if (type == ZB_RLE) { if (n - pos < 1) return ZMT_ST_TRUNCATED; B.comp_size = 1; B.regen_hint = bs; pos += 1; }
        else {
            if (n - pos < bs || bs < 2) return ZMT_ST_TRUNCATED;
            const uint8_t* s = frame + pos;
            const uint32_t b0 = s[0], lt = b0 & 3, sf = (b0 >> 2) & 3;
            uint32_t lregen, lcomp, lhdr;
            if (lt < 2) {
                if (sf == 0 || sf == 2) { lregen = b0 >> 3; lhdr = 1; }
                else if (sf == 1) { if (bs < 2) return ZMT_ST_BLOCK; lregen = (b0 >> 4) | ((uint32_t)s[1] << 4); lhdr = 2; }
                else { if (bs < 3) return ZMT_ST_BLOCK; lregen = (b0 >> 4) | ((uint32_t)s[1] << 4) | ((uint32_t)s[2] << 12); lhdr = 3; }
                lcomp = lt == 0 ? lregen : 1;
            } else {
                if (bs < 5) return ZMT_ST_BLOCK;
                if (sf < 2) { const uint32_t v = s[0] | (s[1] << 8) | ((uint32_t)s[2] << 16); lregen = (v >> 4) & 0x3FF; lcomp = (v >> 14) & 0x3FF; lhdr = 3; }
                else if (sf == 2) { const uint32_t v = h_rd32(s); lregen = (v >> 4) & 0x3FFF; lcomp = (v >> 18) & 0x3FFF; lhdr = 4; }
                else { const uint64_t v = (uint64_t)h_rd32(s) | ((uint64_t)s[4] << 32); lregen = (uint32_t)((v >> 4) & 0x3FFFF); lcomp = (uint32_t)((v >> 22) & 0x3FFFF); lhdr = 5; }
            }
            if ((uint64_t)lhdr + lcomp + 1 > bs) return ZMT_ST_BLOCK;
            const uint8_t* q = s + lhdr + lcomp; const uint32_t qn = bs - lhdr - lcomp;
            uint32_t nseq; const uint32_t q0 = q[0];
            if (q0 == 0) nseq = 0; else if (q0 < 128) nseq = q0;
            else if (q0 < 255) { if (qn < 2) return ZMT_ST_BLOCK; nseq = ((q0 - 128) << 8) + q[1]; }
            else { if (qn < 3) return ZMT_ST_BLOCK; nseq = q[1] + (q[2] << 8) + 0x7F00; }
            B.comp_size = bs; B.regen_hint = lregen; B.nseq = nseq;
            B.seq_off = *scratch_used; *scratch_used += (((uint64_t)nseq * sizeof(ZDSeq)) + 15) & ~15ull;
            B.lit_off = *scratch_used; *scratch_used += ((uint64_t)lregen + 15) & ~15ull;
            pos += bs;
        }
        (*nblocks_io)++;
        first = false;
        if (last) break;
    }
    return pos == n ? ZMT_ST_OK : ZMT_ST_TRAILING;
}

extern "C" size_t zmt_zstd_blk_desc_bytes(void) { return sizeof(ZBlk); }

// workspace: [regen u32 x nblocks][done u32 x nblocks][blk_out u64 x nblocks][expect u64 x nframes][scratch ...]
extern "C" size_t zmt_zstdd_workspace_bytes(uint32_t nframes, uint32_t nblocks, uint64_t scratch_bytes)
{
    return (size_t)(((uint64_t)nblocks * 4 + 255) & ~255ull) * 2 + (((uint64_t)nblocks * 8 + 255) & ~255ull) + (((uint64_t)nframes * 8 + 255) & ~255ull)
           + scratch_bytes + 1024;
}

// d_blocks: nblocks descriptors (device copy of what zmt_zstd_scan_frame_host produced); d_frame_first_blk: nframes+1;
// d_expect: content size per frame (from the frame headers)
extern "C" int zmt_zstd_decompress_device(const void* d_in, const void* d_blocks, uint32_t nblocks, const uint32_t* d_frame_first_blk,
                                          const uint64_t* d_expect, uint32_t nframes, void* d_out, const uint64_t* d_out_off,
                                          uint64_t* d_out_size, uint32_t* d_status, void* d_work, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nframes == 0) return ZMT_ST_OK;
    const int ti = zd_tables_init(); if (ti != ZMT_ST_OK) return ti;
    uint8_t* w = (uint8_t*)d_work;
    uint32_t* regen = (uint32_t*)w; w += (((uint64_t)nblocks * 4 + 255) & ~255ull);
    uint32_t* done = (uint32_t*)w; w += (((uint64_t)nblocks * 4 + 255) & ~255ull);
    uint64_t* blk_out = (uint64_t*)w; w += (((uint64_t)nblocks * 8 + 255) & ~255ull);
    w += (((uint64_t)nframes * 8 + 255) & ~255ull);
    uint8_t* scratch = w;
    cudaMemsetAsync(d_status, 0, (size_t)nframes * 4, stream);
    cudaMemsetAsync(regen, 0, (size_t)nblocks * 4, stream);
    cudaMemsetAsync(done, 0, (size_t)nblocks * 4, stream);
    if (nblocks) {
        zstd_entropy_kernel<<<(nblocks + ZD_WARPS - 1) / ZD_WARPS, 32 * ZD_WARPS, 0, stream>>>((const uint8_t*)d_in, (const ZBlk*)d_blocks, nblocks, scratch, regen, d_status);
    }
    zstd_offsets_kernel<<<(nframes + 127) / 128, 128, 0, stream>>>((const ZBlk*)d_blocks, nblocks, d_frame_first_blk, regen, blk_out, d_out_off, d_expect,
                                                                   (unsigned long long*)d_out_size, d_status, nframes);
    if (nblocks) {
        zstd_execute_kernel<<<(nblocks + ZX_WARPS - 1) / ZX_WARPS, 32 * ZX_WARPS, 0, stream>>>((const uint8_t*)d_in, (const ZBlk*)d_blocks, nblocks, scratch, regen, blk_out,
                                                                                           d_out_off, (uint8_t*)d_out, done, d_status);
    }
    return cudaGetLastError() == cudaSuccess ? ZMT_ST_OK : ZMT_ST_CUDA;
}
