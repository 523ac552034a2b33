/* gmp.h -- declarations-only stand-in for the GNU MP header (GMP 6.x ABI, libgmp.so.10).
 *
 * TEST INFRASTRUCTURE ONLY.  The image ships the GMP runtime
 * (/usr/lib/x86_64-linux-gnu/libgmp.so.10) but not its development header.  This file
 * declares exactly the types, macros and entry points that the reference library's own
 * sources use, so that oracle/Makefile can compile those sources where they lie
 * under /root/reference and link them against the system runtime.  Nothing here is an
 * implementation: every function resolves to the exported __gmp* symbol of libgmp.so.10.
 * Type layouts follow the GMP 6 ABI on LP64 (x86-64): limb = unsigned long (64 bit).
 */
#ifndef PBC_B200_GMP_SHIM_H
#define PBC_B200_GMP_SHIM_H

#include <stddef.h>
#include <stdio.h>
#include <stdarg.h>

#ifdef __cplusplus
extern "C" {
#endif

#define __GNU_MP_VERSION 6
#define __GNU_MP_VERSION_MINOR 3
#define __GNU_MP_VERSION_PATCHLEVEL 0

typedef unsigned long mp_limb_t;
typedef long mp_limb_signed_t;
typedef unsigned long mp_bitcnt_t;
typedef long mp_size_t;
typedef long mp_exp_t;
typedef mp_limb_t *mp_ptr;
typedef const mp_limb_t *mp_srcptr;

#define GMP_LIMB_BITS 64
#define GMP_NAIL_BITS 0
#define GMP_NUMB_BITS 64
#define mp_bits_per_limb 64

typedef struct { int _mp_alloc; int _mp_size; mp_limb_t *_mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

typedef struct { __mpz_struct _mp_num; __mpz_struct _mp_den; } __mpq_struct;
typedef __mpq_struct mpq_t[1];
typedef __mpq_struct *mpq_ptr;
typedef const __mpq_struct *mpq_srcptr;

typedef struct { int _mp_prec; int _mp_size; mp_exp_t _mp_exp; mp_limb_t *_mp_d; } __mpf_struct;
typedef __mpf_struct mpf_t[1];
typedef __mpf_struct *mpf_ptr;
typedef const __mpf_struct *mpf_srcptr;

typedef enum { GMP_RAND_ALG_DEFAULT = 0, GMP_RAND_ALG_LC = 0 } gmp_randalg_t;
typedef struct {
  mpz_t _mp_seed;
  gmp_randalg_t _mp_alg;
  union { void *_mp_lc; } _mp_algdata;
} __gmp_randstate_struct;
typedef __gmp_randstate_struct gmp_randstate_t[1];
typedef __gmp_randstate_struct *gmp_randstate_ptr;

/* ---- macros that are macros in the real header too ---- */
#define mpz_sgn(z) ((z)->_mp_size < 0 ? -1 : (z)->_mp_size > 0)
#define mpf_sgn(f) ((f)->_mp_size < 0 ? -1 : (f)->_mp_size > 0)
#define mpz_odd_p(z) (((z)->_mp_size != 0) & (int)((z)->_mp_d[0] & 1))
#define mpz_even_p(z) (!mpz_odd_p(z))
#define mpq_numref(q) (&((q)->_mp_num))
#define mpq_denref(q) (&((q)->_mp_den))

/* ---- memory hook ---- */
void __gmp_set_memory_functions(void *(*)(size_t), void *(*)(void *, size_t, size_t),
                                void (*)(void *, size_t));
#define mp_set_memory_functions __gmp_set_memory_functions

/* ---- mpz ---- */
#define PBC_SHIM_Z(name) __gmpz_##name
#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_init_set __gmpz_init_set
#define mpz_set __gmpz_set
#define mpz_set_ui __gmpz_set_ui
#define mpz_set_si __gmpz_set_si
#define mpz_set_str __gmpz_set_str
#define mpz_set_f __gmpz_set_f
#define mpz_get_ui __gmpz_get_ui
#define mpz_get_si __gmpz_get_si
#define mpz_getlimbn __gmpz_getlimbn
#define mpz_size __gmpz_size
#define mpz_sizeinbase __gmpz_sizeinbase
#define mpz_add __gmpz_add
#define mpz_add_ui __gmpz_add_ui
#define mpz_sub __gmpz_sub
#define mpz_sub_ui __gmpz_sub_ui
#define mpz_mul __gmpz_mul
#define mpz_mul_ui __gmpz_mul_ui
#define mpz_mul_si __gmpz_mul_si
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_neg __gmpz_neg
#define mpz_mod __gmpz_mod
#define mpz_mod_ui __gmpz_fdiv_r_ui
#define mpz_fdiv_r_ui __gmpz_fdiv_r_ui
#define mpz_fdiv_qr __gmpz_fdiv_qr
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
#define mpz_tdiv_q __gmpz_tdiv_q
#define mpz_tdiv_q_2exp __gmpz_tdiv_q_2exp
#define mpz_divexact __gmpz_divexact
#define mpz_divexact_ui __gmpz_divexact_ui
#define mpz_divisible_p __gmpz_divisible_p
#define mpz_pow_ui __gmpz_pow_ui
#define mpz_ui_pow_ui __gmpz_ui_pow_ui
#define mpz_powm __gmpz_powm
#define mpz_powm_ui __gmpz_powm_ui
#define mpz_invert __gmpz_invert
#define mpz_gcd __gmpz_gcd
#define mpz_sqrt __gmpz_sqrt
#define mpz_fac_ui __gmpz_fac_ui
#define mpz_jacobi __gmpz_jacobi
#define mpz_legendre __gmpz_jacobi
#define mpz_probab_prime_p __gmpz_probab_prime_p
#define mpz_nextprime __gmpz_nextprime
#define mpz_perfect_square_p __gmpz_perfect_square_p
#define mpz_perfect_power_p __gmpz_perfect_power_p
#define mpz_cmp __gmpz_cmp
#define mpz_cmp_ui __gmpz_cmp_ui
#define mpz_cmp_si __gmpz_cmp_si
#define mpz_cmpabs_ui __gmpz_cmpabs_ui
#define mpz_setbit __gmpz_setbit
#define mpz_tstbit __gmpz_tstbit
#define mpz_scan1 __gmpz_scan1
#define mpz_popcount __gmpz_popcount
#define mpz_fits_ulong_p __gmpz_fits_ulong_p
#define mpz_import __gmpz_import
#define mpz_export __gmpz_export
#define mpz_out_str __gmpz_out_str
#define mpz_out_raw __gmpz_out_raw
#define mpz_urandomm __gmpz_urandomm
#define mpz_get_str __gmpz_get_str
#define _mpz_realloc __gmpz_realloc

void mpz_init(mpz_ptr);
void mpz_clear(mpz_ptr);
void mpz_init_set(mpz_ptr, mpz_srcptr);
void mpz_set(mpz_ptr, mpz_srcptr);
void mpz_set_ui(mpz_ptr, unsigned long);
void mpz_set_si(mpz_ptr, long);
int mpz_set_str(mpz_ptr, const char *, int);
void mpz_set_f(mpz_ptr, mpf_srcptr);
unsigned long mpz_get_ui(mpz_srcptr);
long mpz_get_si(mpz_srcptr);
mp_limb_t mpz_getlimbn(mpz_srcptr, mp_size_t);
size_t mpz_size(mpz_srcptr);
size_t mpz_sizeinbase(mpz_srcptr, int);
void mpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_sub_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_mul_si(mpz_ptr, mpz_srcptr, long);
void mpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_neg(mpz_ptr, mpz_srcptr);
void mpz_mod(mpz_ptr, mpz_srcptr, mpz_srcptr);
unsigned long mpz_fdiv_r_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_fdiv_qr(mpz_ptr, mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_tdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_tdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_divexact(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_divexact_ui(mpz_ptr, mpz_srcptr, unsigned long);
int mpz_divisible_p(mpz_srcptr, mpz_srcptr);
void mpz_pow_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_ui_pow_ui(mpz_ptr, unsigned long, unsigned long);
void mpz_powm(mpz_ptr, mpz_srcptr, mpz_srcptr, mpz_srcptr);
void mpz_powm_ui(mpz_ptr, mpz_srcptr, unsigned long, mpz_srcptr);
int mpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_gcd(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_sqrt(mpz_ptr, mpz_srcptr);
void mpz_fac_ui(mpz_ptr, unsigned long);
int mpz_jacobi(mpz_srcptr, mpz_srcptr);
int mpz_probab_prime_p(mpz_srcptr, int);
void mpz_nextprime(mpz_ptr, mpz_srcptr);
int mpz_perfect_square_p(mpz_srcptr);
int mpz_perfect_power_p(mpz_srcptr);
int mpz_cmp(mpz_srcptr, mpz_srcptr);
int mpz_cmp_ui(mpz_srcptr, unsigned long);
int mpz_cmp_si(mpz_srcptr, long);
int mpz_cmpabs_ui(mpz_srcptr, unsigned long);
void mpz_setbit(mpz_ptr, mp_bitcnt_t);
int mpz_tstbit(mpz_srcptr, mp_bitcnt_t);
mp_bitcnt_t mpz_scan1(mpz_srcptr, mp_bitcnt_t);
mp_bitcnt_t mpz_popcount(mpz_srcptr);
int mpz_fits_ulong_p(mpz_srcptr);
void mpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void *mpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
size_t mpz_out_str(FILE *, int, mpz_srcptr);
size_t mpz_out_raw(FILE *, mpz_srcptr);
void mpz_urandomm(mpz_ptr, gmp_randstate_t, mpz_srcptr);
char *mpz_get_str(char *, int, mpz_srcptr);
void *_mpz_realloc(mpz_ptr, mp_size_t);

/* ---- mpn ---- */
#define mpn_add_n __gmpn_add_n
#define mpn_sub_n __gmpn_sub_n
#define mpn_add_1 __gmpn_add_1
#define mpn_sub_1 __gmpn_sub_1
#define mpn_cmp __gmpn_cmp
#define mpn_mul_1 __gmpn_mul_1
#define mpn_addmul_1 __gmpn_addmul_1
#define mpn_mul_n __gmpn_mul_n
#define mpn_sqr __gmpn_sqr
#define mpn_sqr_n(d, s, n) __gmpn_sqr((d), (s), (n))
#define mpn_lshift __gmpn_lshift
#define mpn_rshift __gmpn_rshift
#define mpn_tdiv_qr __gmpn_tdiv_qr

mp_limb_t mpn_add_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_sub_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_add_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t mpn_sub_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
int mpn_cmp(mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_mul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t mpn_addmul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
void mpn_mul_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
void __gmpn_sqr(mp_ptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_lshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_limb_t mpn_rshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
void mpn_tdiv_qr(mp_ptr, mp_ptr, mp_size_t, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);

/* ---- mpq ---- */
#define mpq_init __gmpq_init
#define mpq_clear __gmpq_clear
#define mpq_canonicalize __gmpq_canonicalize
#define mpq_set_ui __gmpq_set_ui
#define mpq_add __gmpq_add
#define mpq_sub __gmpq_sub
#define mpq_inv __gmpq_inv
void mpq_init(mpq_ptr);
void mpq_clear(mpq_ptr);
void mpq_canonicalize(mpq_ptr);
void mpq_set_ui(mpq_ptr, unsigned long, unsigned long);
void mpq_add(mpq_ptr, mpq_srcptr, mpq_srcptr);
void mpq_sub(mpq_ptr, mpq_srcptr, mpq_srcptr);
void mpq_inv(mpq_ptr, mpq_srcptr);

/* ---- mpf (curve-generation code only; never on the pairing path) ---- */
#define mpf_init __gmpf_init
#define mpf_init2 __gmpf_init2
#define mpf_clear __gmpf_clear
#define mpf_set __gmpf_set
#define mpf_set_ui __gmpf_set_ui
#define mpf_set_d __gmpf_set_d
#define mpf_set_q __gmpf_set_q
#define mpf_set_z __gmpf_set_z
#define mpf_set_default_prec __gmpf_set_default_prec
#define mpf_add __gmpf_add
#define mpf_add_ui __gmpf_add_ui
#define mpf_sub __gmpf_sub
#define mpf_mul __gmpf_mul
#define mpf_mul_ui __gmpf_mul_ui
#define mpf_mul_2exp __gmpf_mul_2exp
#define mpf_div __gmpf_div
#define mpf_div_ui __gmpf_div_ui
#define mpf_div_2exp __gmpf_div_2exp
#define mpf_ui_div __gmpf_ui_div
#define mpf_neg __gmpf_neg
#define mpf_cmp __gmpf_cmp
#define mpf_sqrt __gmpf_sqrt
#define mpf_sqrt_ui __gmpf_sqrt_ui
#define mpf_pow_ui __gmpf_pow_ui
#define mpf_get_ui __gmpf_get_ui
#define mpf_get_d __gmpf_get_d
#define mpf_trunc __gmpf_trunc
#define mpf_out_str __gmpf_out_str
void mpf_init(mpf_ptr);
void mpf_init2(mpf_ptr, mp_bitcnt_t);
void mpf_clear(mpf_ptr);
void mpf_set(mpf_ptr, mpf_srcptr);
void mpf_set_ui(mpf_ptr, unsigned long);
void mpf_set_d(mpf_ptr, double);
void mpf_set_q(mpf_ptr, mpq_srcptr);
void mpf_set_z(mpf_ptr, mpz_srcptr);
void mpf_set_default_prec(mp_bitcnt_t);
void mpf_add(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_add_ui(mpf_ptr, mpf_srcptr, unsigned long);
void mpf_sub(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_mul(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_mul_ui(mpf_ptr, mpf_srcptr, unsigned long);
void mpf_mul_2exp(mpf_ptr, mpf_srcptr, mp_bitcnt_t);
void mpf_div(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_div_ui(mpf_ptr, mpf_srcptr, unsigned long);
void mpf_div_2exp(mpf_ptr, mpf_srcptr, mp_bitcnt_t);
void mpf_ui_div(mpf_ptr, unsigned long, mpf_srcptr);
void mpf_neg(mpf_ptr, mpf_srcptr);
int mpf_cmp(mpf_srcptr, mpf_srcptr);
void mpf_sqrt(mpf_ptr, mpf_srcptr);
void mpf_sqrt_ui(mpf_ptr, unsigned long);
void mpf_pow_ui(mpf_ptr, mpf_srcptr, unsigned long);
unsigned long mpf_get_ui(mpf_srcptr);
double mpf_get_d(mpf_srcptr);
void mpf_trunc(mpf_ptr, mpf_srcptr);
size_t mpf_out_str(FILE *, int, size_t, mpf_srcptr);

/* ---- random state + formatted output ---- */
#define gmp_randinit_default __gmp_randinit_default
#define gmp_randseed_ui __gmp_randseed_ui
#define gmp_randclear __gmp_randclear
#define gmp_printf __gmp_printf
#define gmp_fprintf __gmp_fprintf
#define gmp_sprintf __gmp_sprintf
#define gmp_snprintf __gmp_snprintf
#define gmp_vsnprintf __gmp_vsnprintf
#define gmp_vfprintf __gmp_vfprintf
#define gmp_vprintf __gmp_vprintf
void gmp_randinit_default(gmp_randstate_t);
void gmp_randseed_ui(gmp_randstate_t, unsigned long);
void gmp_randclear(gmp_randstate_t);
int gmp_printf(const char *, ...);
int gmp_fprintf(FILE *, const char *, ...);
int gmp_sprintf(char *, const char *, ...);
int gmp_snprintf(char *, size_t, const char *, ...);
int gmp_vsnprintf(char *, size_t, const char *, va_list);
int gmp_vfprintf(FILE *, const char *, va_list);
int gmp_vprintf(const char *, va_list);

#ifdef __cplusplus
}
#endif
#endif
