// stubs.cu -- formats not built yet in this tree state report UNSUPPORTED (never a CPU fallback).
#include "common.cuh"
namespace b2 {
#ifndef HAVE_CSR5
Format *make_csr5(const b200spmv_options &) { return nullptr; }
#endif
}
