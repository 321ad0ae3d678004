/* rb_oracle.c -- TEST INFRASTRUCTURE ONLY: CPU oracle for the rigid-body impulse/friction path.
 *
 * Instantiates rb_oracle_body.h (the restatement, with reference file:line citations) for
 * double (_f64, the reference's precision) and float (_f32, checker for the fp32 GPU mode).
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off [-fopenmp] -shared -fPIC).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the result.
 *
 * Pinning: the reference ships no tests of its own (tests/test_simulation.py is 0 bytes), so this restatement is
 * pinned against outputs of the UNMODIFIED reference run in the build container under the fake MuJoCo backend
 * (oracle/make_golden.py -> tests/golden/*.json; checked by tests/test_oracle_golden.py), which in turn reproduce
 * the two published plots that match the shipped code.  Parity against the real MuJoCo C library is unverified.
 */
#include <math.h>
#include <stdlib.h>

static double absr_f64(double x) { return fabs(x); }
static double sqrtr_f64(double x) { return sqrt(x); }
static float absr_f32(float x) { return fabsf(x); }
static float sqrtr_f32(float x) { return sqrtf(x); }

#define REAL double
#define SUF(name) name##_f64
#include "rb_oracle_body.h"
#undef REAL
#undef SUF

#define REAL float
#define SUF(name) name##_f32
#include "rb_oracle_body.h"
#undef REAL
#undef SUF

int rbo_version(void) { return 1; }
#ifdef _OPENMP
#include <omp.h>
int rbo_max_threads(void) { return omp_get_max_threads(); }
void rbo_set_threads(int n) { omp_set_num_threads(n); }
#else
int rbo_max_threads(void) { return 1; }
void rbo_set_threads(int n) { (void)n; }
#endif
