/*
 * klujax_cuda.h — C-ABI of the B200 (sm_100a) twin of klujax's numeric hot path.
 *
 * Every entry point replaces one XLA-FFI handler of the reference
 * (/root/reference/klujax.cpp; capsule export at :1386-1429, Python
 * registration at /root/reference/klujax.py:834-1015).  The shapes are the
 * canonical ones the reference's Python layer produces before binding
 * (klujax.py:315-351, :1748-1849):
 *     Ai, Aj : int32 [n_nz]              COO row / column indices
 *     Ax     : T     [n_lhs, n_nz]       values, system-major
 *     b, x   : T     [n_lhs, n_col, n_rhs]
 * T = double (…_f64) or interleaved (re,im) double pairs (…_c128).
 *
 * Conventions
 *  - "_dev" pointers are device memory on the current CUDA device, "_host"
 *    pointers are host memory.  Numeric entry points only ENQUEUE work on
 *    `stream` (a cudaStream_t passed as void*); they never synchronise, except
 *    the first numeric call on a symbolic handle, which copies one seed system
 *    to the host to fix the pivot sequence (see DESIGN.md "seeding").
 *  - Return value: 0 = enqueued / done, otherwise a KLU status code (< 0) for
 *    argument errors detected on the host; klujax_cuda_last_error() has text.
 *  - Numerical failure is reported PER SYSTEM, on the device, KLU-style:
 *    status_dev[m] = 0 (KLU_OK), 1 (KLU_SINGULAR: zero or non-finite pivot),
 *    2 (KLUJAX_CUDA_PIVOT_DRIFT: the pivot-parity certificate failed, see below),
 *    -3 (KLU_INVALID: index out of range / not in the analysed pattern);
 *    growth_dev[m] = the certificate of system m: max |L(i,k)| over the columns
 *    whose static pivot is the diagonal candidate, and for a column whose static
 *    pivot is off-diagonal (KLU took the column maximum at seeding time)
 *    max |L(i,k)| / tol and |L(diag,k)| / tol^2 - KLU's own threshold pivoting
 *    (tol = 1e-3) would have chosen the same pivots while it stays <= 1/tol =
 *    1000 (ties aside).  Either pointer may be NULL.  The reference collapses
 *    this into one ffi::Error for the whole batch (klujax.cpp:311-314); a
 *    binding reproduces that by raising if any status != 0.
 *  - Handles are opaque non-zero uint64 values, 0 = null, exactly like the
 *    reference's pointer-as-u64 handles (klujax.cpp:395, :690, :1372).
 *    A numeric handle array has one element per system; elements of one
 *    `factor` call share device storage that is released when the last
 *    element is freed (klujax.cpp:1282-1288 frees element by element).
 */
#ifndef KLUJAX_CUDA_H
#define KLUJAX_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KLUJAX_CUDA_ABI_VERSION 1

int klujax_cuda_abi_version(void);
const char* klujax_cuda_last_error(void);

/* ---- analyze / free_symbolic : klujax.cpp:1331-1374, :1293-1315 ---------- */
int klujax_cuda_analyze(int n_col, int n_nz, const int32_t* Ai_host, const int32_t* Aj_host,
                        uint64_t* symbolic_out);
/* *freed_out: 1 = freed, 0 = nothing to free (null handle), like the i32 result */
int klujax_cuda_free_symbolic(uint64_t symbolic, int32_t* freed_out);

/* ---- solve_f64 / solve_c128 : klujax.cpp:251-380 ------------------------- */
/* analyze + factor + solve + free in one call (indices on the host because the
 * analysis runs there). */
int klujax_cuda_solve_f64(int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai_host,
                          const int32_t* Aj_host, const double* Ax_dev, const double* b_dev,
                          double* x_dev, int32_t* status_dev, double* growth_dev, void* stream);
int klujax_cuda_solve_c128(int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai_host,
                           const int32_t* Aj_host, const double* Ax_dev, const double* b_dev,
                           double* x_dev, int32_t* status_dev, double* growth_dev, void* stream);

/* ---- solve_with_symbol_* : klujax.cpp:382-509 ; tsolve_with_symbol_* : :511-636
 * Ai/Aj are DEVICE arrays here (they arrive as XLA buffers on every call and
 * may be ordered differently from the analyze call, klujax.py:319-321).
 * Ax is [n_lhs, n_nz] and b is [n_lhs, n_col, n_rhs] here; the _ex entry points below take a
 * single matrix or a single right-hand side without materialising it (klujax.py:1844-1848). */
int klujax_cuda_solve_with_symbol_f64(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz,
                                      const int32_t* Ai_dev, const int32_t* Aj_dev,
                                      const double* Ax_dev, const double* b_dev, double* x_dev,
                                      int32_t* status_dev, double* growth_dev, void* stream);
int klujax_cuda_solve_with_symbol_c128(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz,
                                       const int32_t* Ai_dev, const int32_t* Aj_dev,
                                       const double* Ax_dev, const double* b_dev, double* x_dev,
                                       int32_t* status_dev, double* growth_dev, void* stream);
int klujax_cuda_tsolve_with_symbol_f64(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz,
                                       const int32_t* Ai_dev, const int32_t* Aj_dev,
                                       const double* Ax_dev, const double* b_dev, double* x_dev,
                                       int32_t* status_dev, double* growth_dev, void* stream);
int klujax_cuda_tsolve_with_symbol_c128(uint64_t symbolic, int n_lhs, int n_col, int n_rhs,
                                        int n_nz, const int32_t* Ai_dev, const int32_t* Aj_dev,
                                        const double* Ax_dev, const double* b_dev, double* x_dev,
                                        int32_t* status_dev, double* growth_dev, void* stream);

/* The entry points an XLA-FFI shim binds for solve_with_symbol / tsolve_with_symbol
 * (INTEGRATION.md).  n_lhs_A and n_lhs_b are n_lhs or 1 (broadcast, stride 0 on the device).
 * check = 0: enqueue only, per-system status and growth are left in the arrays.
 * check != 0 reproduces the reference's semantics (klujax.cpp:311-314, :441): the stream is
 * synchronised, systems whose pivot-parity certificate failed (status 2) or that met an exactly
 * zero static pivot (status 1) are solved again on the GPU with a pivot sequence seeded from one
 * of THEM, and the return value is 0 or the status of the first system that still fails
 * (*first_bad_out = its index, *n_redone_out = systems solved again).  status 2 =
 * KLUJAX_CUDA_PIVOT_DRIFT: KLU's threshold pivoting would not have kept the static pivots. */
#define KLUJAX_CUDA_PIVOT_DRIFT 2
int klujax_cuda_solve_with_symbol_ex_f64(uint64_t symbolic, int n_lhs, int n_lhs_A, int n_lhs_b,
                                         int n_col, int n_rhs, int n_nz, const int32_t* Ai_dev,
                                         const int32_t* Aj_dev, const double* Ax_dev,
                                         const double* b_dev, double* x_dev, int32_t* status_dev,
                                         double* growth_dev, int check, int32_t* first_bad_out,
                                         int32_t* n_redone_out, void* stream);
int klujax_cuda_solve_with_symbol_ex_c128(uint64_t symbolic, int n_lhs, int n_lhs_A, int n_lhs_b,
                                          int n_col, int n_rhs, int n_nz, const int32_t* Ai_dev,
                                          const int32_t* Aj_dev, const double* Ax_dev,
                                          const double* b_dev, double* x_dev, int32_t* status_dev,
                                          double* growth_dev, int check, int32_t* first_bad_out,
                                          int32_t* n_redone_out, void* stream);
int klujax_cuda_tsolve_with_symbol_ex_f64(uint64_t symbolic, int n_lhs, int n_lhs_A, int n_lhs_b,
                                          int n_col, int n_rhs, int n_nz, const int32_t* Ai_dev,
                                          const int32_t* Aj_dev, const double* Ax_dev,
                                          const double* b_dev, double* x_dev, int32_t* status_dev,
                                          double* growth_dev, int check, int32_t* first_bad_out,
                                          int32_t* n_redone_out, void* stream);
int klujax_cuda_tsolve_with_symbol_ex_c128(uint64_t symbolic, int n_lhs, int n_lhs_A, int n_lhs_b,
                                           int n_col, int n_rhs, int n_nz, const int32_t* Ai_dev,
                                           const int32_t* Aj_dev, const double* Ax_dev,
                                           const double* b_dev, double* x_dev, int32_t* status_dev,
                                           double* growth_dev, int check, int32_t* first_bad_out,
                                           int32_t* n_redone_out, void* stream);

/* The check + re-seed half of the _ex entry points, for a batch whose solves have been enqueued
 * already (e.g. a host batch streamed in chunks on several streams, all joined into `stream`). */
int klujax_cuda_reseed_failed_f64(int transpose, int n_lhs, int n_lhs_A, int n_lhs_b, int n_col,
                                  int n_rhs, int n_nz, const int32_t* Ai_dev, const int32_t* Aj_dev,
                                  const double* Ax_dev, const double* b_dev, double* x_dev,
                                  int32_t* status_dev, double* growth_dev, int32_t* first_bad_out,
                                  int32_t* n_redone_out, void* stream);
int klujax_cuda_reseed_failed_c128(int transpose, int n_lhs, int n_lhs_A, int n_lhs_b, int n_col,
                                   int n_rhs, int n_nz, const int32_t* Ai_dev, const int32_t* Aj_dev,
                                   const double* Ax_dev, const double* b_dev, double* x_dev,
                                   int32_t* status_dev, double* growth_dev, int32_t* first_bad_out,
                                   int32_t* n_redone_out, void* stream);

/* ---- factor_* : klujax.cpp:638-731 --------------------------------------- */
/* numeric_out_host[n_lhs] receives the handles (host array: see INTEGRATION.md). */
int klujax_cuda_factor_f64(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai_dev,
                           const int32_t* Aj_dev, const double* Ax_dev, uint64_t* numeric_out_host,
                           int32_t* status_dev, double* growth_dev, void* stream);
int klujax_cuda_factor_c128(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai_dev,
                            const int32_t* Aj_dev, const double* Ax_dev, uint64_t* numeric_out_host,
                            int32_t* status_dev, double* growth_dev, void* stream);

/* The entry a shim binds for `factor`: check != 0 restores the reference's semantics (klu_factor
 * pivots every system, klujax.cpp:679): the stream is synchronised and systems whose pivot-parity
 * certificate failed are factorized again in a child batch with a pivot sequence seeded from one
 * of them; the handles returned may then point into several batches (every entry point accepts
 * such mixed arrays).  On failure everything is freed (klujax.cpp:684-687), the status of the
 * first failing system is returned and *first_bad_out is its index. */
int klujax_cuda_factor_ex_f64(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai_dev,
                              const int32_t* Aj_dev, const double* Ax_dev, uint64_t* numeric_out_host,
                              int32_t* status_dev, double* growth_dev, int check,
                              int32_t* first_bad_out, void* stream);
int klujax_cuda_factor_ex_c128(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai_dev,
                               const int32_t* Aj_dev, const double* Ax_dev, uint64_t* numeric_out_host,
                               int32_t* status_dev, double* growth_dev, int check,
                               int32_t* first_bad_out, void* stream);

/* ---- refactor_* : klujax.cpp:733-833 (in place; output handles = input handles).  klu_refactor
 * keeps the pivot order of the handle and never re-pivots: status is 0 or 1 here, growth is
 * reported for information only. */
int klujax_cuda_refactor_f64(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai_dev,
                             const int32_t* Aj_dev, const double* Ax_dev,
                             const uint64_t* numeric_in_host, uint64_t* numeric_out_host,
                             int32_t* status_dev, double* growth_dev, void* stream);
int klujax_cuda_refactor_c128(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai_dev,
                              const int32_t* Aj_dev, const double* Ax_dev,
                              const uint64_t* numeric_in_host, uint64_t* numeric_out_host,
                              int32_t* status_dev, double* growth_dev, void* stream);

/* ---- refactor_and_solve_* : klujax.cpp:835-993 --------------------------- */
int klujax_cuda_refactor_and_solve_f64(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz,
                                       const int32_t* Ai_dev, const int32_t* Aj_dev,
                                       const double* Ax_dev, const double* b_dev,
                                       const uint64_t* numeric_in_host, double* x_dev,
                                       uint64_t* numeric_out_host, int32_t* status_dev,
                                       double* growth_dev, void* stream);
int klujax_cuda_refactor_and_solve_c128(uint64_t symbolic, int n_lhs, int n_col, int n_rhs,
                                        int n_nz, const int32_t* Ai_dev, const int32_t* Aj_dev,
                                        const double* Ax_dev, const double* b_dev,
                                        const uint64_t* numeric_in_host, double* x_dev,
                                        uint64_t* numeric_out_host, int32_t* status_dev,
                                        double* growth_dev, void* stream);

/* ---- solve_with_numeric_* : klujax.cpp:995-1137 ; tsolve_with_numeric_* : :1139-1271
 * b is canonical [n_lhs_b, n_col, n_rhs]; n_numeric or n_lhs_b may be 1 and is
 * then broadcast (klujax.cpp:1044-1059); x is [max(n_numeric, n_lhs_b), n_col, n_rhs]. */
int klujax_cuda_solve_with_numeric_f64(uint64_t symbolic, int n_numeric,
                                       const uint64_t* numeric_host, int n_lhs_b, int n_col,
                                       int n_rhs, const double* b_dev, double* x_dev, void* stream);
int klujax_cuda_solve_with_numeric_c128(uint64_t symbolic, int n_numeric,
                                        const uint64_t* numeric_host, int n_lhs_b, int n_col,
                                        int n_rhs, const double* b_dev, double* x_dev, void* stream);
int klujax_cuda_tsolve_with_numeric_f64(uint64_t symbolic, int n_numeric,
                                        const uint64_t* numeric_host, int n_lhs_b, int n_col,
                                        int n_rhs, const double* b_dev, double* x_dev, void* stream);
int klujax_cuda_tsolve_with_numeric_c128(uint64_t symbolic, int n_numeric,
                                         const uint64_t* numeric_host, int n_lhs_b, int n_col,
                                         int n_rhs, const double* b_dev, double* x_dev,
                                         void* stream);

/* ---- free_numeric : klujax.cpp:1273-1291 --------------------------------- */
int klujax_cuda_free_numeric(int n, const uint64_t* numeric_host, int32_t* freed_out);

/* ---- the same entry points with DEVICE-resident handle arrays -------------------------------
 * On the CUDA platform XLA delivers `numeric` as a u64[n_lhs] device buffer (klujax.cpp:392-395,
 * :690, :1044 read it on the host).  Output handles are written stream-ordered, without a
 * synchronisation.  Input handles must be known on the host to pick the batch and its schedule:
 * they are copied back, 8 bytes per system and ONE stream synchronisation per call. */
int klujax_cuda_factor_dev_f64(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai_dev,
                               const int32_t* Aj_dev, const double* Ax_dev, uint64_t* numeric_out_dev,
                               int32_t* status_dev, double* growth_dev, int check,
                               int32_t* first_bad_out, void* stream);
int klujax_cuda_factor_dev_c128(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai_dev,
                                const int32_t* Aj_dev, const double* Ax_dev, uint64_t* numeric_out_dev,
                                int32_t* status_dev, double* growth_dev, int check,
                                int32_t* first_bad_out, void* stream);
int klujax_cuda_refactor_dev_f64(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz,
                                 const int32_t* Ai_dev, const int32_t* Aj_dev, const double* Ax_dev,
                                 const double* b_unused, const uint64_t* numeric_in_dev, double* x_unused,
                                 uint64_t* numeric_out_dev, int32_t* status_dev, double* growth_dev,
                                 void* stream);
int klujax_cuda_refactor_dev_c128(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz,
                                  const int32_t* Ai_dev, const int32_t* Aj_dev, const double* Ax_dev,
                                  const double* b_unused, const uint64_t* numeric_in_dev, double* x_unused,
                                  uint64_t* numeric_out_dev, int32_t* status_dev, double* growth_dev,
                                  void* stream);
int klujax_cuda_refactor_and_solve_dev_f64(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz,
                                           const int32_t* Ai_dev, const int32_t* Aj_dev,
                                           const double* Ax_dev, const double* b_dev,
                                           const uint64_t* numeric_in_dev, double* x_dev,
                                           uint64_t* numeric_out_dev, int32_t* status_dev,
                                           double* growth_dev, void* stream);
int klujax_cuda_refactor_and_solve_dev_c128(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz,
                                            const int32_t* Ai_dev, const int32_t* Aj_dev,
                                            const double* Ax_dev, const double* b_dev,
                                            const uint64_t* numeric_in_dev, double* x_dev,
                                            uint64_t* numeric_out_dev, int32_t* status_dev,
                                            double* growth_dev, void* stream);
int klujax_cuda_solve_with_numeric_dev_f64(uint64_t symbolic, int n_numeric, const uint64_t* numeric_dev,
                                           int n_lhs_b, int n_col, int n_rhs, const double* b_dev,
                                           double* x_dev, void* stream);
int klujax_cuda_solve_with_numeric_dev_c128(uint64_t symbolic, int n_numeric, const uint64_t* numeric_dev,
                                            int n_lhs_b, int n_col, int n_rhs, const double* b_dev,
                                            double* x_dev, void* stream);
int klujax_cuda_tsolve_with_numeric_dev_f64(uint64_t symbolic, int n_numeric, const uint64_t* numeric_dev,
                                            int n_lhs_b, int n_col, int n_rhs, const double* b_dev,
                                            double* x_dev, void* stream);
int klujax_cuda_tsolve_with_numeric_dev_c128(uint64_t symbolic, int n_numeric, const uint64_t* numeric_dev,
                                             int n_lhs_b, int n_col, int n_rhs, const double* b_dev,
                                             double* x_dev, void* stream);
int klujax_cuda_free_numeric_dev(int n, const uint64_t* numeric_dev, int32_t* freed_out, void* stream);

/* ---- dot_f64 / dot_c128 : klujax.cpp:165-249 (batched COO SpMM, "next" row f1) */
int klujax_cuda_dot_f64(int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai_dev,
                        const int32_t* Aj_dev, const double* Ax_dev, const double* x_dev,
                        double* b_dev, void* stream);
int klujax_cuda_dot_c128(int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai_dev,
                         const int32_t* Aj_dev, const double* Ax_dev, const double* x_dev,
                         double* b_dev, void* stream);

/* ---- transpose of dot with respect to the values : klujax.py:1695-1716 (dot_transpose:
 * dAx[m,e] = sum_k ct[m, Ai[e], k] * x[m, Aj[e], k], pure jnp in the reference; here one
 * gather kernel, used by the fused value-and-vjp of the Python mirror, "next" row f2) */
int klujax_cuda_dot_transpose_ax_f64(int n_lhs, int n_col, int n_rhs, int n_nz,
                                     const int32_t* Ai_dev, const int32_t* Aj_dev,
                                     const double* ct_dev, const double* x_dev, double* dAx_dev,
                                     void* stream);
int klujax_cuda_dot_transpose_ax_c128(int n_lhs, int n_col, int n_rhs, int n_nz,
                                      const int32_t* Ai_dev, const int32_t* Aj_dev,
                                      const double* ct_dev, const double* x_dev, double* dAx_dev,
                                      void* stream);

/* ---- host-only introspection (used by the parity tests; no GPU needed) ---- */
/* Fixes the pivot sequence from one host-resident system of values given in the
 * COO order of the analyze call (what the first numeric call does internally). */
int klujax_cuda_seed_host(uint64_t symbolic, int is_complex, const double* Ax_host);
/* sizes: out[0]=n out[1]=nblocks out[2]=maxblock out[3]=nzoff out[4]=structural_rank
 *        out[5]=nnz(L) strictly lower  out[6]=nnz(U) strictly upper  out[7]=seeded(0/1)
 *        out[8]=regime (0 shared-memory, 1 level-scheduled)  out[9]=levels of factor+solve
 *        out[10]=multiply-adds of one refactor  out[11]=noffdiag
 *        out[12]=batches, out[13]=32-bit words of the packed factor+solve program (regime 0) */
int klujax_cuda_symbolic_info(uint64_t symbolic, int is_complex, int64_t* out14);
/* P[n], Q[n], R[nblocks+1] of the analysis; any pointer may be NULL */
int klujax_cuda_symbolic_perm(uint64_t symbolic, int32_t* P, int32_t* Q, int32_t* R);
/* after seeding: Pnum[n]; CSC patterns of L and U in pivot order with block-local,
 * ascending row indices (Lp/Up have n+1 entries) */
int klujax_cuda_symbolic_pattern(uint64_t symbolic, int is_complex, int32_t* Pnum, int32_t* Lp,
                                 int32_t* Li, int32_t* Up, int32_t* Ui);

/* ---- register-resident regime: planner self-test (host only, TEST INFRASTRUCTURE) ----------
 * Builds the register-resident program of the fused factor + solve (klujax.cpp:431-451) for a
 * seeded symbolic handle and runs it on the planner's HOST INTERPRETER for ONE system: the
 * parity tests check the schedule against the oracle without a GPU.  No operator above ever
 * calls this; the product path is the generated sm_100a kernel and has no CPU fallback.
 *   Ax_host: values in the COO order of the analyze call; b_host, x_host: [n_col, n_rhs]
 *   stats[16]: 0 value registers, 1 multiply-add instructions, 2 useful multiply-adds,
 *     3 shuffle broadcasts, 4 pivots, 5 multiplier instructions, 6 zero-fill instructions,
 *     7 parking stores, 8 parked-row multiply-adds, 9 planes of parked values, 10 epochs,
 *     11 rows kept in registers, 12 load instructions, 13 instructions in total,
 *     14 16-bit tables, 15 shared memory per warp (bytes)
 *   ptx_path: if not NULL the generated PTX is written there (tests assemble it with ptxas) */
int klujax_cuda_rowreg_selftest(uint64_t symbolic, int is_complex, int transpose, int n_rhs,
                                const double* Ax_host, const double* b_host, double* x_host,
                                int32_t* status_out, double* growth_out, int64_t* stats16,
                                const char* ptx_path);

/* 1 if the library was built with the XLA-FFI handlers of klujax_b200/csrc/klujax_cuda_ffi.cc
 * (xla/ffi/api/ffi.h on the include path), 0 if that file compiled to its stub. */
int klujax_cuda_ffi_available(void);

#ifdef __cplusplus
}
#endif
#endif /* KLUJAX_CUDA_H */
