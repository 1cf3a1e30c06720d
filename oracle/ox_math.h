/* Oracle (TEST INFRASTRUCTURE ONLY): elementary-function selection.
 * Default build: C libm.  -DOX_PMATH: portable algorithms made of individually rounded IEEE operations
 * (ox_pmath.h), the same operation sequences as xnet_b200/csrc/xb_pmath.cuh, so that the CUDA strict mode
 * can be checked bit for bit.  pow(x,y) is exp(y*log(x)) in the portable build. */
#ifndef OX_MATH_H
#define OX_MATH_H
#include <math.h>
#ifdef OX_PMATH
#include "ox_pmath.h"
#define M_EXP(x) oxp_exp(x)
#define M_LOG(x) oxp_log(x)
#define M_LOG10(x) oxp_log10(x)
#define M_POW10(x) oxp_pow10(x)
#define M_POW(x, y) oxp_exp((y) * oxp_log(x))
#else
#define M_EXP(x) exp(x)
#define M_LOG(x) log(x)
#define M_LOG10(x) log10(x)
#define M_POW10(x) pow(10.0, (x))
#define M_POW(x, y) pow((x), (y))
#endif
#endif
