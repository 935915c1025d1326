// standalone harness (no torch / python): compress a generated stream with the device API, print sizes + checksum
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "../include/zstdmt_b200_dev.h"
int main(int argc, char** argv)
{
    size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 1048576; uint32_t chunk = 1 << 20; int first = argc > 2 ? atoi(argv[2]) : 1;
    std::vector<uint8_t> h(n + 16); if (argc > 4) { FILE* f = fopen(argv[4], "rb"); } 
    zmt_gen_stream(2, first, 1, chunk, h.data(), n, 4);
    size_t skip = argc > 3 ? strtoull(argv[3], 0, 10) : 0; n -= skip;
    uint32_t nch = zmt_chunk_count(n, chunk);
    uint8_t *d_in, *d_work, *d_out; uint64_t* d_off;
    cudaMalloc(&d_in, n + 256); cudaMalloc(&d_work, zmt_lz4c_workspace_bytes(nch, chunk)); cudaMalloc(&d_out, zmt_lz4c_out_bound(nch, chunk)); cudaMalloc(&d_off, (nch + 1) * 8);
    cudaMemcpy(d_in, h.data() + skip, n, cudaMemcpyHostToDevice);
    int rc = zmt_lz4_compress_device(d_in, n, chunk, nullptr, nch, d_work, d_out, d_off, 0);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint64_t> off(nch + 1); cudaMemcpy(off.data(), d_off, (nch + 1) * 8, cudaMemcpyDeviceToHost);
    printf("rc=%d cuda=%s total=%llu\n", rc, cudaGetErrorString(e), (unsigned long long)off[nch]);
    return 0;
}
