// Placeholder until the specialised thread-per-codeword kernel for the shipped 32x64
// code lands: reports "not this code" so dispatch stays on the generic kernels.
#pragma once
#include "code_tables.h"
#include "decode_kernels.cuh"

namespace {
inline bool tables_match_c4(const ldpc535::CodeTables &) { return false; }
}
namespace ldpc535 {
inline cudaError_t launch_c4_thread(const DecodeParams &, bool, int, cudaStream_t)
{
    return cudaErrorNotSupported;
}
}
