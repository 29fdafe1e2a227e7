// Nei-Gojobori site and difference tables, built on the host from first
// principles at library load.
//
// Replaces the reference's literal hash tables: get_number_of_sites
// (snpgenie_within_group.pl:1625-1721) and return_avg_diffs (:1728-17996).
// Rules recovered from those tables:
//   sites  per codon position, the fraction of single-nucleotide changes that
//          do not produce a STOP which are nonsynonymous (N) / synonymous (S);
//          STOP codons have 0/0; a codon's sites are position 1+2+3 summed in
//          that order in double precision.
//   diffs  for every ordered codon pair, nonsynonymous/synonymous differences
//          averaged over all mutational pathways that never pass through a
//          STOP; any pair with a STOP endpoint is (0,0); every average is the
//          15-significant-digit decimal the reference's source text holds.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cmath>

#include "snpg_internal.h"

namespace snpg {

namespace {

// amino acid of codon index 16*b0+4*b1+b2 (A,C,G,T = 0..3); '*' = STOP
const char kAminoAcid[65] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";

inline int base_at(int codon, int pos) { return (codon >> (4 - 2 * pos)) & 3; }
inline int with_base(int codon, int pos, int b) { return (codon & ~(3 << (4 - 2 * pos))) | (b << (4 - 2 * pos)); }
inline bool is_stop(int codon) { return kAminoAcid[codon] == '*'; }

struct PathTally { int n_sum = 0, s_sum = 0, paths = 0; };

// Depth-first walk over the orders in which the still-differing positions can
// be changed; a branch dies when it lands on a STOP before reaching the target.
void walk(int cur, int target, int n, int s, PathTally &t) {
    if (cur == target) { t.n_sum += n; t.s_sum += s; t.paths++; return; }
    for (int p = 0; p < 3; p++) {
        if (base_at(cur, p) == base_at(target, p)) continue;
        int nxt = with_base(cur, p, base_at(target, p));
        if (is_stop(nxt)) continue;
        bool syn = kAminoAcid[nxt] == kAminoAcid[cur];
        walk(nxt, target, n + (syn ? 0 : 1), s + (syn ? 1 : 0), t);
    }
}

double as_written(double x) {  // value of the 15-digit decimal literal
    char buf[48];
    snprintf(buf, sizeof buf, "%.15g", x);
    return strtod(buf, nullptr);
}

}  // namespace

void build_tables(double *sites /*[64][2]*/, double *diffs /*[2][64][64]*/) {
    for (int c = 0; c < 64; c++) {
        double N = 0.0, S = 0.0;
        for (int p = 0; p < 3 && !is_stop(c); p++) {
            int syn = 0, non = 0;
            for (int b = 0; b < 4; b++) {
                if (b == base_at(c, p)) continue;
                int m = with_base(c, p, b);
                if (is_stop(m)) continue;
                (kAminoAcid[m] == kAminoAcid[c] ? syn : non)++;
            }
            if (syn + non == 0) continue;
            N += double(non) / double(syn + non);
            S += double(syn) / double(syn + non);
        }
        sites[2 * c] = N;
        sites[2 * c + 1] = S;
    }
    memset(diffs, 0, sizeof(double) * 2 * 64 * 64);
    for (int a = 0; a < 64; a++) {
        if (is_stop(a)) continue;
        for (int b = 0; b < 64; b++) {
            if (b == a || is_stop(b)) continue;
            PathTally t;
            walk(a, b, 0, 0, t);
            if (!t.paths) continue;
            diffs[a * 64 + b] = as_written(double(t.n_sum) / t.paths);
            diffs[4096 + a * 64 + b] = as_written(double(t.s_sum) / t.paths);
        }
    }
}

}  // namespace snpg
