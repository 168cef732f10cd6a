// common.cuh -- shared definitions of the rsrec CUDA library (sm_100a).
//
// Device data layout ("RI36"): a block vector psi(18,18,kk) of the reference (recursion.f90:66-68) is held as
// (kk+1) site blocks of 648 doubles; column c of a site occupies 36 consecutive doubles, real parts of rows 0..17
// first, imaginary parts next:  re(k,c) = blk[c*36 + k],  im(k,c) = blk[c*36 + 18 + k].  Block kk is the "null
// site": always zero, target of nn(i,j)=0 entries, so gathers need no branch.  A Hamiltonian block H(18,18) is
// held as its 36x36 REAL embedding ("HR36", row-major, 1296 doubles):  Hreal = [[Hr,-Hi],[Hi,Hr]], so that a
// complex 18x18x18 product is the real product  [Cr;Ci](36 x n) = Hreal (36x36) * [Pr;Pi] (36 x n), which is what
// the FP64 tensor-core (DMMA m8n8k4) kernels consume without any sign/offset fix-ups in the inner loop.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NB 18
#define BLKC 324      // complex elements per block
#define BLKD 648      // doubles per block
#define COLD 36       // doubles per RI36 column / HR36 row
#define HBLK 1296     // doubles per HR36 Hamiltonian block (36x36 real embedding)

enum Epilogue {
  EPI_STORE = 0,  // out = acc
  EPI_HAM = 1,    // out = (acc - b*in)/a                      ham_vec_matmul, recursion.f90:974-976
  EPI_CHEB = 2,   // out = 2*((acc - b*in)/a) - prev; grams     chebyshev_recur_ll, recursion.f90:2557-2592
  EPI_CHEB_NOGRAM = 3,  // same without the two reductions      compute_moments_stochastic, recursion.f90:1164
  EPI_HOP = 4,    // out(pmn) = acc - pmn; A += in^H acc        hop_b, recursion.f90:1638-1647
  EPI_HOP_GRAM = 5  // tensor pipeline: EPI_HOP with A = sum in^H (H in) reduced inside the SpMV kernel (no hpsi vector)
};

struct GatherTerm {
  const double *H;    // [ncls][nslot_h][HBLK]
  const double *src;  // block vector(s) gathered through the neighbour table
  int first_slot;     // 0 = include the on-site slot, 1 = neighbours only
};

struct ApplyParams {
  int kk, nslot_h, ngather;
  size_t vstride;          // doubles between consecutive units of a batched vector = (kk+1)*648
  const int32_t *nbr;      // [ngather][kk]  0-based neighbour, kk = null site; row 0 = self
  const int32_t *cls;      // [kk] H class: type-1 for bulk sites, ntype + site for site-indexed (hall) sites
  GatherTerm g[2];
  int ngterms;
  const double *Hx;        // optional on-site extra term Hx[cls] * srcx[self]   (hoh: enim + lsham)
  const double *srcx;
  const double *addend;    // optional + addend[self]                             (hoh: h psi)
  const double *in;        // the vector the epilogue calls "in" (psi1 / psi)
  const double *prev;      // EPI_CHEB: psi0;  EPI_HOP: pmn (== out)
  double *out;
  double *out2;            // EPI_HOP (tensor pipeline): second output = H in (hpsi) next to out = H in - prev
  double a, b;
  int epi;
  double *part;            // gram partials [unit][cta][2][648] (EPI_CHEB: D1,D2; EPI_HOP: A,unused)
};
