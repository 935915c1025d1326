/* zstdmt_b200_harness.h — test / bench harness libraries (NOT part of the product library).
 *
 *   libzmt_datagen.so  deterministic synthetic inputs (zstdmt_b200/harness/datagen.c); no dependencies,
 *                      loaded by both arms of bench.py
 *   libzmt_memio.so    memory-to-memory drivers of the callback API (zstdmt_b200/harness/memio_glue.c), linked against
 *                      libzstdmt_b200.so; the same source is compiled with -DGLUE_PREFIX=ref_ into oracle/_ref so both
 *                      arms run byte-identical fn_read / fn_write callbacks
 */
#ifndef ZSTDMT_B200_HARNESS_H
#define ZSTDMT_B200_HARNESS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic inputs (harness/datagen.c) ---- */
void zmt_gen_chunk(int kind, uint64_t chunk_index, uint8_t* buf, size_t n);
void zmt_gen_stream(int kind, uint64_t first, uint64_t stride, size_t chunk, uint8_t* buf, size_t total, int nthreads);
/* share of consumer `rank` of `world` when the global stream is dealt round-robin in batches of `batch` chunks:
 * local chunk c = global chunk (c / batch) * batch * world + rank * batch + c % batch */
void zmt_gen_stream_dealt(int kind, uint64_t rank, uint64_t world, uint64_t batch, size_t chunk, uint8_t* buf, size_t total, int nthreads);

/* ---- memory-to-memory drivers of the callback API (harness/memio_glue.c) ----
 * stats[0..4] = bytes written, frames, Insize counter, Outsize counter, (reads<<32 | writes) */
size_t zmt_lz4_compress_mem(int threads, int level, int chunk, const void* src, size_t n, void* dst, size_t cap, size_t* stats);
size_t zmt_lz4_decompress_mem(int threads, int inputsize, const void* src, size_t n, void* dst, size_t cap, size_t* stats);
size_t zmt_zstd_compress_mem(int threads, int level, int chunk, const void* src, size_t n, void* dst, size_t cap, size_t* stats);
size_t zmt_zstd_decompress_mem(int threads, int inputsize, const void* src, size_t n, void* dst, size_t cap, size_t* stats);

#ifdef __cplusplus
}
#endif
#endif
