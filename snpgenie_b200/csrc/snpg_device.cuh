// Device helpers shared by kernels.cu and tile_hist.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace snpg {

constexpr unsigned kFull = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------
// Residue classification.  Normalisation follows the reference ingest: uc, then
// U->T (snpgenie_within_group.pl:231-232); on the '-' strand the residue is
// first complemented with tr/ACGT/TGCA/ exactly as fasta2revcom.pl:77-79 does
// (so U is NOT complemented there and then becomes T).
// class: 0..3 = A,C,G,T; 4 = undefined (N or '-'; wg:493,498); 5 = other.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint8_t norm_char(uint8_t c, int minus) {
    if (c >= 'a' && c <= 'z') c -= 32;
    if (minus) {
        if (c == 'A') c = 'T';
        else if (c == 'C') c = 'G';
        else if (c == 'G') c = 'C';
        else if (c == 'T') c = 'A';
    }
    if (c == 'U') c = 'T';
    return c;
}
__device__ __forceinline__ uint8_t class_of(uint8_t n) {
    return n == 'A' ? 0 : n == 'C' ? 1 : n == 'G' ? 2 : n == 'T' ? 3 : (n == 'N' || n == '-') ? 4 : 5;
}

}  // namespace snpg
