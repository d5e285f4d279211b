/*
 * TEST INFRASTRUCTURE ONLY -- stands in for the header the reference's CMake
 * generates with FortranCInterface (reference CMakeLists.txt:156-168): maps the
 * upper-case BLAS/LAPACK names used in src/linear_algebra/*.h onto the
 * gfortran-style lower-case-underscore symbols an LP64 BLAS exports.
 */
#ifndef MGMOL_B200_ORACLE_FC_MANGLE_H
#define MGMOL_B200_ORACLE_FC_MANGLE_H
#define DAXPY daxpy_
#define DCOPY dcopy_
#define DDOT ddot_
#define DGEMM dgemm_
#define DGEMV dgemv_
#define DLANGE dlange_
#define DNRM2 dnrm2_
#define DPOCON dpocon_
#define DPOTRF dpotrf_
#define DPOTRI dpotri_
#define DPOTRS dpotrs_
#define DROT drot_
#define DSCAL dscal_
#define DSYEV dsyev_
#define DSYGST dsygst_
#define DSYGV dsygv_
#define DSYMM dsymm_
#define DSYMV dsymv_
#define DSYRK dsyrk_
#define DTRMM dtrmm_
#define DTRSM dtrsm_
#define DTRTRS dtrtrs_
#define IDAMAX idamax_
#define ISAMAX isamax_
#define SAXPY saxpy_
#define SCOPY scopy_
#define SDOT sdot_
#define SGEMM sgemm_
#define SGEMV sgemv_
#define SNRM2 snrm2_
#define SROT srot_
#define SSCAL sscal_
#define SSYRK ssyrk_
#define STRSM strsm_
#define DGETRF dgetrf_
#define DGETRS dgetrs_
#define SSYMM ssymm_
#define STRMM strmm_
#endif
