// calc.hpp -- arbitrary-precision expression evaluator for the command line of the B200 ECM
// driver.  Same input language as the reference's calc.c (calc_init/calc/calc_finalize,
// calc.c:1106-1126): + - * / % ^ << >> ! # parentheses, unary minus, decimal / 0x hex literals and
// the functions fib luc gcd jacobi sqrt modinv modexp nroot shift xor and or not abs lg2 rand randb
// lte gte.  Precedence as in calc.c:389-402 (shifts < additive < multiplicative < power).
#pragma once
#include <string>
#include "gmp.h"

// Evaluates expr into result (must be mpz_init'ed).  Returns an empty string on success, otherwise
// a description of the syntax error.
std::string calc_eval(const std::string &expr, mpz_t result);
