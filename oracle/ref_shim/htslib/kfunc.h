/*  ORACLE-ONLY stub of <htslib/kfunc.h>; only referenced by test16() which is unreachable without -a PV4. */
#ifndef ORACLE_STUB_KFUNC_H
#define ORACLE_STUB_KFUNC_H
double kf_erfc(double x);
double kf_betai(double a, double b, double x);
double kt_fisher_exact(int n11, int n12, int n21, int n22, double *_left, double *_right, double *two);
#endif
