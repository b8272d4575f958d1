#include "expr.h"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace eucl {

namespace {

struct Token {
    enum Kind { Num, Ident, Op, LParen, RParen, Comma } kind;
    double num = 0;
    std::string text;
};

struct FuncDesc {
    const char* name;
    int func;
    int arity;
};

const FuncDesc FUNCS[] = {
    {"sqrt", EUCL_FN_SQRT, 1},   {"abs", EUCL_FN_ABS, 1},     {"exp", EUCL_FN_EXP, 1},
    {"ln", EUCL_FN_LN, 1},       {"sin", EUCL_FN_SIN, 1},     {"cos", EUCL_FN_COS, 1},
    {"tan", EUCL_FN_TAN, 1},     {"asin", EUCL_FN_ASIN, 1},   {"acos", EUCL_FN_ACOS, 1},
    {"atan", EUCL_FN_ATAN, 1},   {"sinh", EUCL_FN_SINH, 1},   {"cosh", EUCL_FN_COSH, 1},
    {"tanh", EUCL_FN_TANH, 1},   {"floor", EUCL_FN_FLOOR, 1}, {"ceil", EUCL_FN_CEIL, 1},
    {"round", EUCL_FN_ROUND, 1}, {"signum", EUCL_FN_SIGNUM, 1},
    {"atan2", EUCL_FN_ATAN2, 2}, {"max", EUCL_FN_MAX, 2},     {"min", EUCL_FN_MIN, 2},
};

const FuncDesc* find_func(const std::string& name) {
    for (const auto& f : FUNCS)
        if (name == f.name) return &f;
    return nullptr;
}

bool tokenize(const std::string& s, std::vector<Token>* out, std::string* error) {
    size_t i = 0;
    while (i < s.size()) {
        char c = s[i];
        if (std::isspace((unsigned char)c)) {
            ++i;
            continue;
        }
        if (std::isdigit((unsigned char)c) || (c == '.' && i + 1 < s.size() && std::isdigit((unsigned char)s[i + 1]))) {
            const char* begin = s.c_str() + i;
            char* end = nullptr;
            double v = std::strtod(begin, &end);
            if (end == begin) {
                *error = "bad number";
                return false;
            }
            Token t;
            t.kind = Token::Num;
            t.num = v;
            out->push_back(t);
            i += (size_t)(end - begin);
            continue;
        }
        if (std::isalpha((unsigned char)c) || c == '_') {
            size_t j = i;
            while (j < s.size() && (std::isalnum((unsigned char)s[j]) || s[j] == '_')) ++j;
            Token t;
            t.kind = Token::Ident;
            t.text = s.substr(i, j - i);
            out->push_back(t);
            i = j;
            continue;
        }
        Token t;
        if (c == '(') t.kind = Token::LParen;
        else if (c == ')') t.kind = Token::RParen;
        else if (c == ',') t.kind = Token::Comma;
        else if (std::strchr("+-*/%^", c)) {
            t.kind = Token::Op;
            t.text = std::string(1, c);
        } else {
            *error = std::string("unexpected character '") + c + "'";
            return false;
        }
        out->push_back(t);
        ++i;
    }
    return true;
}

// operator-stack entries
struct StackOp {
    enum Kind { Binary, Unary, Func, Paren } kind;
    char op = 0;
    const FuncDesc* func = nullptr;
    int argc = 0; // Func: commas seen + 1
};

int precedence(const StackOp& o) {
    if (o.kind == StackOp::Unary) return 3;
    if (o.kind == StackOp::Binary) {
        switch (o.op) {
        case '+':
        case '-': return 1;
        case '*':
        case '/':
        case '%': return 2;
        case '^': return 4;
        }
    }
    return 0;
}

void emit(const StackOp& o, std::vector<EuclExprOp>* out) {
    EuclExprOp e{};
    if (o.kind == StackOp::Unary) {
        if (o.op == '+') return; // unary plus is a no-op
        e.op = EUCL_EX_NEG;
    } else if (o.kind == StackOp::Binary) {
        switch (o.op) {
        case '+': e.op = EUCL_EX_ADD; break;
        case '-': e.op = EUCL_EX_SUB; break;
        case '*': e.op = EUCL_EX_MUL; break;
        case '/': e.op = EUCL_EX_DIV; break;
        case '%': e.op = EUCL_EX_REM; break;
        case '^': e.op = EUCL_EX_POW; break;
        }
    } else if (o.kind == StackOp::Func) {
        e.op = o.func->arity == 1 ? EUCL_EX_FUNC1 : EUCL_EX_FUNC2;
        e.arg = o.func->func;
    }
    out->push_back(e);
}

} // namespace

bool expr_compile(const std::string& text, const std::string& legend, int dim,
                  std::vector<EuclExprOp>* out, std::string* error) {
    std::vector<Token> toks;
    if (!tokenize(text, &toks, error)) return false;
    if (toks.empty()) {
        *error = "empty expression";
        return false;
    }
    std::vector<StackOp> stack;
    std::vector<EuclExprOp> rpn;
    bool expect_operand = true; // true at start / after an operator, '(' or ','
    for (size_t i = 0; i < toks.size(); ++i) {
        const Token& t = toks[i];
        switch (t.kind) {
        case Token::Num: {
            if (!expect_operand) {
                *error = "unexpected number";
                return false;
            }
            EuclExprOp e{};
            e.op = EUCL_EX_CONST;
            e.value = t.num;
            rpn.push_back(e);
            expect_operand = false;
            break;
        }
        case Token::Ident: {
            if (!expect_operand) {
                *error = "unexpected identifier `" + t.text + "`";
                return false;
            }
            bool is_call = i + 1 < toks.size() && toks[i + 1].kind == Token::LParen;
            if (is_call) {
                const FuncDesc* f = find_func(t.text);
                if (!f) {
                    *error = "unknown function `" + t.text + "`";
                    return false;
                }
                StackOp o;
                o.kind = StackOp::Func;
                o.func = f;
                o.argc = 1;
                stack.push_back(o);
                // the '(' that follows is pushed by the LParen case
                break;
            }
            EuclExprOp e{};
            size_t pos = t.text.size() == 1 ? legend.find(t.text[0]) : std::string::npos;
            if (pos != std::string::npos && (int)pos < dim) {
                e.op = EUCL_EX_VAR;
                e.arg = (int)pos;
            } else if (t.text == "pi") {
                e.op = EUCL_EX_CONST;
                e.value = 3.14159265358979323846;
            } else if (t.text == "e") {
                e.op = EUCL_EX_CONST;
                e.value = 2.71828182845904523536;
            } else {
                *error = "unknown variable `" + t.text + "` (legend `" + legend + "`)";
                return false;
            }
            rpn.push_back(e);
            expect_operand = false;
            break;
        }
        case Token::Op: {
            StackOp o;
            o.op = t.text[0];
            if (expect_operand) {
                if (o.op != '-' && o.op != '+') {
                    *error = "unexpected operator `" + t.text + "`";
                    return false;
                }
                o.kind = StackOp::Unary;
                stack.push_back(o);
                break;
            }
            o.kind = StackOp::Binary;
            int p = precedence(o);
            bool right_assoc = o.op == '^';
            while (!stack.empty()) {
                const StackOp& top = stack.back();
                if (top.kind != StackOp::Binary && top.kind != StackOp::Unary) break;
                int tp = precedence(top);
                if (tp > p || (tp == p && !right_assoc)) {
                    emit(top, &rpn);
                    stack.pop_back();
                } else {
                    break;
                }
            }
            stack.push_back(o);
            expect_operand = true;
            break;
        }
        case Token::LParen: {
            if (!expect_operand) {
                *error = "unexpected '('";
                return false;
            }
            StackOp o;
            o.kind = StackOp::Paren;
            stack.push_back(o);
            break;
        }
        case Token::Comma: {
            if (expect_operand) {
                *error = "unexpected ','";
                return false;
            }
            while (!stack.empty() && stack.back().kind != StackOp::Paren) {
                emit(stack.back(), &rpn);
                stack.pop_back();
            }
            if (stack.size() < 2 || stack[stack.size() - 2].kind != StackOp::Func) {
                *error = "',' outside a function call";
                return false;
            }
            stack[stack.size() - 2].argc++;
            expect_operand = true;
            break;
        }
        case Token::RParen: {
            if (expect_operand) {
                *error = "unexpected ')'";
                return false;
            }
            while (!stack.empty() && stack.back().kind != StackOp::Paren) {
                emit(stack.back(), &rpn);
                stack.pop_back();
            }
            if (stack.empty()) {
                *error = "unbalanced ')'";
                return false;
            }
            stack.pop_back(); // the paren
            if (!stack.empty() && stack.back().kind == StackOp::Func) {
                if (stack.back().argc != stack.back().func->arity) {
                    *error = std::string("wrong number of arguments for `") + stack.back().func->name + "`";
                    return false;
                }
                emit(stack.back(), &rpn);
                stack.pop_back();
            }
            expect_operand = false;
            break;
        }
        }
    }
    if (expect_operand) {
        *error = "expression ends with an operator";
        return false;
    }
    while (!stack.empty()) {
        if (stack.back().kind == StackOp::Paren || stack.back().kind == StackOp::Func) {
            *error = "unbalanced '('";
            return false;
        }
        emit(stack.back(), &rpn);
        stack.pop_back();
    }
    if (rpn.size() > 64) {
        *error = "expression too long";
        return false;
    }
    out->insert(out->end(), rpn.begin(), rpn.end());
    return true;
}

double expr_eval(const EuclExprOp* ops, int len, const double* vars) {
    double st[64];
    int sp = 0;
    for (int i = 0; i < len; ++i) {
        const EuclExprOp& o = ops[i];
        switch (o.op) {
        case EUCL_EX_CONST: st[sp++] = o.value; break;
        case EUCL_EX_VAR: st[sp++] = vars[o.arg]; break;
        case EUCL_EX_NEG: st[sp - 1] = -st[sp - 1]; break;
        case EUCL_EX_FUNC1: {
            double x = st[sp - 1], r = x;
            switch (o.arg) {
            case EUCL_FN_SQRT: r = std::sqrt(x); break;
            case EUCL_FN_ABS: r = std::fabs(x); break;
            case EUCL_FN_EXP: r = std::exp(x); break;
            case EUCL_FN_LN: r = std::log(x); break;
            case EUCL_FN_SIN: r = std::sin(x); break;
            case EUCL_FN_COS: r = std::cos(x); break;
            case EUCL_FN_TAN: r = std::tan(x); break;
            case EUCL_FN_ASIN: r = std::asin(x); break;
            case EUCL_FN_ACOS: r = std::acos(x); break;
            case EUCL_FN_ATAN: r = std::atan(x); break;
            case EUCL_FN_SINH: r = std::sinh(x); break;
            case EUCL_FN_COSH: r = std::cosh(x); break;
            case EUCL_FN_TANH: r = std::tanh(x); break;
            case EUCL_FN_FLOOR: r = std::floor(x); break;
            case EUCL_FN_CEIL: r = std::ceil(x); break;
            case EUCL_FN_ROUND: r = std::round(x); break;
            case EUCL_FN_SIGNUM: r = std::isnan(x) ? x : (std::signbit(x) ? -1.0 : 1.0); break;
            }
            st[sp - 1] = r;
            break;
        }
        default: {
            double b = st[--sp], a = st[sp - 1], r = 0;
            switch (o.op) {
            case EUCL_EX_ADD: r = a + b; break;
            case EUCL_EX_SUB: r = a - b; break;
            case EUCL_EX_MUL: r = a * b; break;
            case EUCL_EX_DIV: r = a / b; break;
            case EUCL_EX_REM: r = std::fmod(a, b); break;
            case EUCL_EX_POW: r = std::pow(a, b); break;
            case EUCL_EX_FUNC2:
                switch (o.arg) {
                case EUCL_FN_ATAN2: r = std::atan2(a, b); break;
                case EUCL_FN_MAX: r = std::fmax(a, b); break;
                case EUCL_FN_MIN: r = std::fmin(a, b); break;
                }
                break;
            }
            st[sp - 1] = r;
        }
        }
    }
    return sp > 0 ? st[sp - 1] : 0.0;
}

} // namespace eucl
