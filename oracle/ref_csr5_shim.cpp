// ref_csr5_shim.cpp -- flat C entry point around the reference's UNMODIFIED CSR5 conversion routines.
//
// TEST INFRASTRUCTURE ONLY (see oracle/spmv_oracle.c).  The reference vendors Liu & Vinter's CSR5
// benchmark under opt/Benchmark_SpMV_using_CSR5/; its CUDA variant cannot be built here, its AVX2
// variant implements the same conversion (CSR5_avx2/detail/avx2/format_avx2.h) with the tile width as
// a macro.  This file includes those headers where they lie (-I$(REF)/opt/.../CSR5_avx2) and overrides
// ANONYMOUSLIB_CSR5_OMEGA to 32 -- the GPU tile width, CSR5_cuda/detail/cuda/common_cuda.h:11 -- after
// common_avx2.h has been seen once (its include guard keeps the override in force inside format_avx2.h).
// The driver steps below restate CSR5_avx2/anonymouslib_avx2.h:105-216 (asCSR5), which only sizes the
// arrays and calls the three conversion routines.
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <iostream>
#include <omp.h>
#include <mm_malloc.h>
#include <sys/time.h>
using namespace std;
#include "detail/avx2/common_avx2.h"
#undef ANONYMOUSLIB_CSR5_OMEGA
#define ANONYMOUSLIB_CSR5_OMEGA 32
#include "detail/avx2/utils_avx2.h"
#include "detail/avx2/format_avx2.h"

extern "C" __attribute__((visibility("default")))
int ref_csr5_shape(int nnz, int sigma, int *bit_y, int *bit_ss, int *num_packet, int *p)
{
    int base = 2, by = 1, bs = 1;                                   // anonymouslib_avx2.h:121-128
    while (base < ANONYMOUSLIB_CSR5_OMEGA * sigma) { base *= 2; by++; }
    base = 2;
    while (base < ANONYMOUSLIB_CSR5_OMEGA) { base *= 2; bs++; }
    int bit_all = by + bs + sigma;
    *bit_y = by; *bit_ss = bs;
    *num_packet = (int)ceil((double)bit_all / 32.0);                 // :133-134
    *p = (int)ceil((double)nnz / (double)(ANONYMOUSLIB_CSR5_OMEGA * sigma));   // :137
    return 0;
}

// col/val are converted in place.  offset must hold at least nnz + m entries (upper bound).
extern "C" __attribute__((visibility("default")))
int ref_csr5_convert(int m, int nnz, int sigma, const int *row_ptr_in, int *col, double *val,
                     unsigned *tile_ptr, unsigned *desc, int *offset_ptr, int *offset, int *num_offsets_out)
{
    omp_set_num_threads(1);        // the reference's s2 has a benign `+=` race across threads (format_avx2.h:203)
    int by, bs, num_packet, p;
    ref_csr5_shape(nnz, sigma, &by, &bs, &num_packet, &p);
    // format_avx2.h:48-55 reads row_pointer[m+1] when a tile's span ends at row m: give it a sentinel
    int *row_ptr = (int *)malloc(sizeof(int) * ((size_t)m + 2));
    memcpy(row_ptr, row_ptr_in, sizeof(int) * ((size_t)m + 1));
    row_ptr[m + 1] = -1;
    memset(desc, 0, sizeof(unsigned) * (size_t)p * ANONYMOUSLIB_CSR5_OMEGA * num_packet);
    memset(offset_ptr, 0, sizeof(int) * ((size_t)p + 1));
    int err = generate_partition_pointer<int, unsigned>(sigma, p, m, nnz, tile_ptr, row_ptr);
    int num_offsets = 0;
    if (!err)
        err = generate_partition_descriptor<int, unsigned>(sigma, p, m, by, bs, num_packet, row_ptr, tile_ptr, desc,
                                                           offset_ptr, &num_offsets);
    if (!err && num_offsets)
        err = generate_partition_descriptor_offset<int, unsigned>(sigma, p, by, bs, num_packet, row_ptr, tile_ptr, desc,
                                                                  offset_ptr, offset);
    if (!err) err = aosoa_transpose<int, unsigned, double>(sigma, nnz, tile_ptr, col, val, true);
    *num_offsets_out = num_offsets;
    free(row_ptr);
    return err;
}
