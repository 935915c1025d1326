/* zstdmt_b200_dev.h — device-resident batch entry points of libzstdmt_b200.so.
 *
 * These are what the host pipeline behind LZ4MT_* / ZSTDCB_* calls per batch, exported so a
 * caller that already holds its data in HBM (bench.py's device-timed metric, a GPU-side
 * producer) can skip the host staging.  Plain C ABI: device pointers + sizes + a CUDA stream
 * handle passed as void*.  All work is enqueued on `stream`; nothing synchronises.
 *
 * What each call replaces in the reference (one call = the codec step of every worker
 * iteration of a whole batch of chunks):
 *   zmt_lz4_compress_device    LZ4F_compressFrame + 12-byte header, lib/lz4-mt_compress.c:280-298
 *   zmt_lz4_decompress_device  LZ4F_decompress, lib/lz4-mt_decompress.c:328-362
 *   zmt_zstd_compress_device   ZSTD_compress + 12-byte header, lib/zstd-mt_compress.c:284-302
 *   zmt_zstd_decompress_device ZSTD_decompressStream loop, lib/zstd-mt_decompress.c:442-527
 */
#ifndef ZSTDMT_B200_DEV_H
#define ZSTDMT_B200_DEV_H
#include <stddef.h>
#include <stdint.h>

#define ZMT_LZ4_TMP_STRIDE 65824u     /* >= LZ4 worst case for a 64 KiB block (65536 + 257 + 16), 16-byte multiple */

/* per-frame / per-call status codes */
#define ZMT_ST_OK               0u
#define ZMT_ST_TRUNCATED        1u
#define ZMT_ST_BAD_MAGIC        2u
#define ZMT_ST_BAD_HEADER       3u
#define ZMT_ST_HDR_CHECKSUM     4u
#define ZMT_ST_BLOCK            5u
#define ZMT_ST_DST_SMALL        6u
#define ZMT_ST_CONTENT_CHECKSUM 7u
#define ZMT_ST_CONTENT_SIZE     8u
#define ZMT_ST_TRAILING         9u
#define ZMT_ST_UNSUPPORTED      10u
#define ZMT_ST_CUDA             11u
#define ZMT_ST_BAD_ARG          12u
#define ZMT_ST_HAS_CHK          0x100u   /* internal flag between decode and verify kernels */

#ifdef __cplusplus
extern "C" {
#endif

/* number of chunks (= frames) an input of in_bytes cut every chunk_size bytes produces;
 * an empty input still yields one frame (lib/lz4-mt_compress.c:265) */
uint32_t zmt_chunk_count(uint64_t in_bytes, uint32_t chunk_size);

/* ---- LZ4 ----
 * Input layout: chunk c starts at d_in + c*chunk_size.  Its length is d_chunk_bytes[c] when
 * that (device) array is given, else derived from in_bytes (all full, last one short).
 * Output: the framed stream, frames back to back; d_frame_off[c] = offset of frame c's
 * 12-byte skippable header, d_frame_off[nchunks] = total bytes. */
size_t   zmt_lz4c_workspace_bytes(uint32_t nchunks, uint32_t chunk_size);
uint64_t zmt_lz4c_out_bound(uint32_t nchunks, uint32_t chunk_size);
int      zmt_lz4_compress_device(const void* d_in, uint64_t in_bytes, uint32_t chunk_size, const uint32_t* d_chunk_bytes,
                                 uint32_t nchunks, void* d_work, void* d_out, uint64_t* d_frame_off, void* stream);

/* d_in/in_bytes = the framed stream in HBM; d_frame_off[f] = offset of frame f's 12-byte header, d_frame_csize[f] =
 * its payload size; d_out_off[f..f+1] = where frame f's content goes and how much room it has;
 * nslots = sum over frames of max(1, ceil(room_f / 64 KiB)): the capacity of the block table (one warp per LZ4F block,
 * two passes: token parse + literals, then match execution; frames with linked blocks — what liblz4 emits for the
 * reference — decode block-parallel too, a block waits on the previous one only where a match reaches into it);
 * d_out_size[f] / d_status[f] receive decoded bytes and a ZMT_ST_* code. */
size_t   zmt_lz4d_workspace_bytes(uint32_t nframes, uint32_t nslots, uint64_t in_bytes);
int      zmt_lz4_decompress_device(const void* d_in, uint64_t in_bytes, const uint64_t* d_frame_off, const uint32_t* d_frame_csize, uint32_t nframes,
                                   uint32_t nslots, void* d_out, const uint64_t* d_out_off, uint64_t* d_out_size,
                                   uint32_t* d_status, void* d_work, void* stream);

/* ---- Zstandard ----  same layout contract as the LZ4 entry points; frames are single-segment zstd frames
 * (magic, FHD, content size, blocks, no checksum) behind the same 12-byte container header. */
size_t   zmt_zstdc_workspace_bytes(uint32_t nchunks, uint32_t chunk_size);
uint64_t zmt_zstdc_out_bound(uint32_t nchunks, uint32_t chunk_size);
int      zmt_zstd_compress_device(const void* d_in, uint64_t in_bytes, uint32_t chunk_size, const uint32_t* d_chunk_bytes,
                                  uint32_t nchunks, void* d_work, void* d_out, uint64_t* d_frame_off, void* stream);

/* Decode.  The host walks frame + block headers (3 bytes per block, plus the two section headers of compressed
 * blocks) with zmt_zstd_scan_frame_host — exactly what the pipeline's reader thread does while it parses frames —
 * and hands the device a block table: one warp per block for entropy decoding, a scan for the output offsets, one
 * warp per block for sequence execution.  Descriptors are opaque (zmt_zstd_blk_desc_bytes() each). */
size_t   zmt_zstd_blk_desc_bytes(void);
int      zmt_zstd_scan_frame_host(const uint8_t* frame, size_t n, uint64_t base_off, uint32_t frame_idx, void* blocks_out,
                                  uint32_t* nblocks_io, uint32_t max_blocks, uint64_t* scratch_used, uint64_t* content_size,
                                  uint32_t* needs_seq /* 1: blocks depend on earlier ones -> frame-sequential entropy pass */);
size_t   zmt_zstdd_workspace_bytes(uint32_t nframes, uint32_t nblocks, uint64_t scratch_bytes);
int      zmt_zstd_decompress_device(const void* d_in, const void* d_blocks, uint32_t nblocks, const uint32_t* d_frame_first_blk,
                                    const uint64_t* d_expect, const uint32_t* d_frame_seq, uint32_t nframes, void* d_out,
                                    const uint64_t* d_out_off, uint64_t* d_out_size, uint32_t* d_status, void* d_work, void* stream);

/* ---- multi-GPU: one LZ4MT_* / ZSTDCB_* call deals its batches round-robin over the devices named by the
 * environment variable ZSTDMT_GPUS ("all" or "0,1,..."; default: the calling thread's current device) and re-serialises
 * the frames on the host.  zmt_device_batches(dev) = batches submitted to that device since the library was loaded. */
uint64_t zmt_device_batches(int dev);

/* ---- per-kernel device timing (CUDA events on the launching stream) ----
 * zmt_prof_begin() arms it; every kernel launched by the entry points above is bracketed by two
 * events; zmt_prof_end() (after the caller synchronised the stream) sums them per kernel id. */
enum { ZMT_K_LZ4_COMPRESS = 0, ZMT_K_XXH32, ZMT_K_LZ4_SIZES, ZMT_K_SCAN, ZMT_K_LZ4_PACK, ZMT_K_LZ4_DECODE, ZMT_K_XXH32_DEC,
       ZMT_K_ZSTD_COMPRESS, ZMT_K_ZSTD_PACK, ZMT_K_ZSTD_DECODE, ZMT_K_LZ4_DEXEC, ZMT_K_COUNT };
/* LZ4 decode: ZMT_K_LZ4_DECODE = pass A (token parse + literals), ZMT_K_LZ4_DEXEC = pass B (match execution) */
void zmt_prof_begin(void);
int  zmt_prof_end(double* ms, int* count, int max_ids);

#ifdef __cplusplus
}
#endif
#endif
