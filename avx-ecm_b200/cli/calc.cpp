// calc.cpp -- precedence-climbing evaluator over GMP integers (see calc.hpp).
#include "calc.hpp"
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace {

struct Val {
    mpz_t v;
    Val() { mpz_init(v); }
    Val(const Val &o) { mpz_init(v); mpz_set(v, o.v); }
    Val &operator=(const Val &o) { mpz_set(v, o.v); return *this; }
    ~Val() { mpz_clear(v); }
};

struct Parser {
    const std::string &s;
    size_t p = 0;
    gmp_randstate_t rs;
    explicit Parser(const std::string &str) : s(str) { gmp_randinit_default(rs); gmp_randseed_ui(rs, 42); }
    ~Parser() { gmp_randclear(rs); }

    [[noreturn]] void fail(const std::string &m) { throw std::runtime_error(m + " at position " + std::to_string(p)); }
    void skip() { while (p < s.size() && isspace((unsigned char)s[p])) p++; }
    bool eat(const char *tok)
    {
        skip();
        size_t n = strlen(tok);
        if (s.compare(p, n, tok) == 0) { p += n; return true; }
        return false;
    }

    // precedence levels: 0 shifts, 1 additive, 2 multiplicative, 3 power (right associative)
    Val expr(int minprec = 0)
    {
        Val lhs = unary();
        for (;;) {
            skip();
            int prec; std::string op;
            if (s.compare(p, 2, "<<") == 0) { prec = 0; op = "<<"; }
            else if (s.compare(p, 2, ">>") == 0) { prec = 0; op = ">>"; }
            else if (p < s.size() && (s[p] == '+' || s[p] == '-')) { prec = 1; op = s.substr(p, 1); }
            else if (p < s.size() && (s[p] == '*' || s[p] == '/' || s[p] == '%')) { prec = 2; op = s.substr(p, 1); }
            else if (p < s.size() && s[p] == '^') { prec = 3; op = "^"; }
            else break;
            if (prec < minprec) break;
            p += op.size();
            Val rhs = expr(op == "^" ? prec : prec + 1);
            apply(op, lhs, rhs);
        }
        return lhs;
    }

    void apply(const std::string &op, Val &a, const Val &b)
    {
        if (op == "+") mpz_add(a.v, a.v, b.v);
        else if (op == "-") mpz_sub(a.v, a.v, b.v);
        else if (op == "*") mpz_mul(a.v, a.v, b.v);
        else if (op == "/") { if (mpz_sgn(b.v) == 0) fail("division by zero"); mpz_tdiv_q(a.v, a.v, b.v); }
        else if (op == "%") { if (mpz_sgn(b.v) == 0) fail("division by zero"); mpz_mod(a.v, a.v, b.v); }
        else if (op == "^") mpz_pow_ui(a.v, a.v, mpz_get_ui(b.v));
        else if (op == "<<") mpz_mul_2exp(a.v, a.v, mpz_get_ui(b.v));
        else if (op == ">>") mpz_tdiv_q_2exp(a.v, a.v, mpz_get_ui(b.v));
    }

    Val unary()
    {
        skip();
        if (eat("-")) { Val v = unary(); mpz_neg(v.v, v.v); return v; }
        if (eat("+")) return unary();
        return postfix();
    }

    Val postfix()
    {
        Val v = primary();
        for (;;) {
            skip();
            if (p < s.size() && s[p] == '!') { p++; mpz_fac_ui(v.v, mpz_get_ui(v.v)); }
            else if (p < s.size() && s[p] == '#') { p++; mpz_primorial_ui(v.v, mpz_get_ui(v.v)); }
            else break;
        }
        return v;
    }

    Val primary()
    {
        skip();
        if (p >= s.size()) fail("unexpected end of expression");
        if (s[p] == '(') {
            p++;
            Val v = expr(0);
            if (!eat(")")) fail("expected ')'");
            return v;
        }
        if (isdigit((unsigned char)s[p])) {
            size_t q = p;
            int base = 10;
            if (s.compare(p, 2, "0x") == 0 || s.compare(p, 2, "0X") == 0) { base = 16; q = p + 2; p = q; while (p < s.size() && isxdigit((unsigned char)s[p])) p++; }
            else while (p < s.size() && isdigit((unsigned char)s[p])) p++;
            Val v;
            if (mpz_set_str(v.v, s.substr(q, p - q).c_str(), base) != 0) fail("bad number");
            return v;
        }
        if (isalpha((unsigned char)s[p])) {
            size_t q = p;
            while (p < s.size() && (isalnum((unsigned char)s[p]) || s[p] == '_')) p++;
            std::string name = s.substr(q, p - q);
            if (!eat("(")) fail("expected '(' after " + name);
            std::vector<Val> a;
            if (!eat(")")) {
                do { a.push_back(expr(0)); } while (eat(","));
                if (!eat(")")) fail("expected ')'");
            }
            return call(name, a);
        }
        fail(std::string("unexpected character '") + s[p] + "'");
    }

    Val call(const std::string &f, std::vector<Val> &a)
    {
        auto need = [&](size_t n) { if (a.size() != n) fail(f + " takes " + std::to_string(n) + " argument(s)"); };
        Val r;
        if (f == "fib") { need(1); mpz_fib_ui(r.v, mpz_get_ui(a[0].v)); }
        else if (f == "luc") { need(1); mpz_lucnum_ui(r.v, mpz_get_ui(a[0].v)); }
        else if (f == "gcd") { need(2); mpz_gcd(r.v, a[0].v, a[1].v); }
        else if (f == "jacobi") { need(2); mpz_set_si(r.v, mpz_jacobi(a[0].v, a[1].v)); }
        else if (f == "sqrt") { need(1); mpz_sqrt(r.v, a[0].v); }
        else if (f == "modinv") { need(2); if (!mpz_invert(r.v, a[0].v, a[1].v)) mpz_set_ui(r.v, 0); }
        else if (f == "modexp") { need(3); mpz_powm(r.v, a[0].v, a[1].v, a[2].v); }
        else if (f == "nroot") { need(2); mpz_root(r.v, a[0].v, mpz_get_ui(a[1].v)); }
        else if (f == "shift") {
            need(2);
            if (mpz_sgn(a[1].v) >= 0) mpz_mul_2exp(r.v, a[0].v, mpz_get_ui(a[1].v));
            else { Val t; mpz_neg(t.v, a[1].v); mpz_tdiv_q_2exp(r.v, a[0].v, mpz_get_ui(t.v)); }
        }
        else if (f == "xor") { need(2); mpz_xor(r.v, a[0].v, a[1].v); }
        else if (f == "and") { need(2); mpz_and(r.v, a[0].v, a[1].v); }
        else if (f == "or") { need(2); mpz_ior(r.v, a[0].v, a[1].v); }
        else if (f == "not") { need(1); mpz_com(r.v, a[0].v); }
        else if (f == "abs") { need(1); mpz_abs(r.v, a[0].v); }
        else if (f == "lg2") { need(1); mpz_set_ui(r.v, mpz_sizeinbase(a[0].v, 2)); }
        else if (f == "lte") { need(2); mpz_set_ui(r.v, mpz_cmp(a[0].v, a[1].v) <= 0); }
        else if (f == "gte") { need(2); mpz_set_ui(r.v, mpz_cmp(a[0].v, a[1].v) >= 0); }
        else if (f == "rand") {                       // random number of that many decimal digits
            need(1);
            Val lim; mpz_set_ui(lim.v, 10); mpz_pow_ui(lim.v, lim.v, mpz_get_ui(a[0].v));
            mpz_urandomm(r.v, rs, lim.v);
        }
        else if (f == "randb") { need(1); mpz_urandomb(r.v, rs, mpz_get_ui(a[0].v)); }
        else fail("unknown function " + f);
        return r;
    }
};

}  // namespace

std::string calc_eval(const std::string &expr, mpz_t result)
{
    try {
        Parser ps(expr);
        Val v = ps.expr(0);
        ps.skip();
        if (ps.p != expr.size()) ps.fail("trailing characters");
        mpz_set(result, v.v);
        return "";
    } catch (const std::exception &e) {
        return e.what();
    }
}
