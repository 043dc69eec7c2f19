/* oracle/perturb/perturb.h -- TEST INFRASTRUCTURE.  Force-included when the C restatement is rebuilt
 * with a deliberately perturbed math library (oracle/Makefile target `sens`), to find the cells whose
 * reference results depend on the last bit of libm (tests/conditioning.py). */
#include <math.h>
double pp_pow(double, double);
double pp_pow_explog(double, double);
double pp_exp(double);
double pp_log(double);
double pp_sin(double);
double pp_acos(double);
double pp_var(double);
#ifdef PERTURB_ALL   /* every libm call of the path at once, and the marked variables */
#define PERTURB_POW
#define PERTURB_EXPLOG
#define PERTURB_TRIG
#define PERTURB_VARS
#endif
#ifdef PERTURB_VARS  /* PP_VAR(x) in splash_oracle.c: net radiation, econ, water density, Ksat_visc moved by an ulp */
#define PP_VAR(x) pp_var(x)
#endif
#ifdef PERTURB_POW
#define pow pp_pow
#endif
#ifdef PERTURB_POWEXPLOG
#define pow pp_pow_explog
#endif
#ifdef PERTURB_TRIG
#define sin pp_sin
#define acos pp_acos
#endif
#ifdef PERTURB_EXPLOG
#define exp pp_exp
#define log pp_log
#endif
