#include "json.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace eucl {

namespace {

// json crate, number.rs: exponent_to_power_f64 -- tables of literals 1e0..1e22 / 1e-0..1e-22,
// powf beyond.  The literals are correctly rounded by the compiler, as rustc would.
double exponent_to_power(int e) {
    static const double POS[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,
                                   1e8,  1e9,  1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                   1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    static const double NEG[23] = {1e-0,  1e-1,  1e-2,  1e-3,  1e-4,  1e-5,  1e-6,  1e-7,
                                   1e-8,  1e-9,  1e-10, 1e-11, 1e-12, 1e-13, 1e-14, 1e-15,
                                   1e-16, 1e-17, 1e-18, 1e-19, 1e-20, 1e-21, 1e-22};
    int index = e < 0 ? -e : e;
    if (index < 23) return e < 0 ? NEG[index] : POS[index];
    return std::pow(10.0, (double)e);
}

struct Reader {
    const std::string& s;
    size_t i = 0;
    std::string err;

    explicit Reader(const std::string& text) : s(text) {}

    void skip_ws() {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
    }
    bool fail(const std::string& what) {
        if (err.empty()) {
            char buf[64];
            std::snprintf(buf, sizeof buf, " at byte %zu", i);
            err = what + buf;
        }
        return false;
    }
    bool literal(const char* word) {
        size_t n = 0;
        while (word[n]) ++n;
        if (s.compare(i, n, word) != 0) return fail("unexpected token");
        i += n;
        return true;
    }

    bool parse_string(std::string* out) {
        if (i >= s.size() || s[i] != '"') return fail("expected string");
        ++i;
        out->clear();
        while (i < s.size()) {
            char c = s[i++];
            if (c == '"') return true;
            if (c == '\\') {
                if (i >= s.size()) break;
                char e = s[i++];
                switch (e) {
                case '"': out->push_back('"'); break;
                case '\\': out->push_back('\\'); break;
                case '/': out->push_back('/'); break;
                case 'b': out->push_back('\b'); break;
                case 'f': out->push_back('\f'); break;
                case 'n': out->push_back('\n'); break;
                case 'r': out->push_back('\r'); break;
                case 't': out->push_back('\t'); break;
                case 'u': {
                    if (i + 4 > s.size()) return fail("bad \\u escape");
                    unsigned cp = (unsigned)std::strtoul(s.substr(i, 4).c_str(), nullptr, 16);
                    i += 4;
                    if (cp < 0x80) {
                        out->push_back((char)cp);
                    } else if (cp < 0x800) {
                        out->push_back((char)(0xC0 | (cp >> 6)));
                        out->push_back((char)(0x80 | (cp & 0x3F)));
                    } else {
                        out->push_back((char)(0xE0 | (cp >> 12)));
                        out->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
                        out->push_back((char)(0x80 | (cp & 0x3F)));
                    }
                    break;
                }
                default: return fail("bad escape");
                }
            } else {
                out->push_back(c);
            }
        }
        return fail("unterminated string");
    }

    // (sign, u64 mantissa, decimal exponent); digits that overflow the mantissa only move the
    // exponent, as the crate's parser does.
    bool parse_number(JsonValue* out) {
        out->kind = JsonValue::Number;
        out->negative = false;
        if (s[i] == '-') {
            out->negative = true;
            ++i;
        }
        if (i >= s.size() || s[i] < '0' || s[i] > '9') return fail("expected digit");
        uint64_t m = 0;
        int e = 0;
        auto push_digit = [&](int d, bool fractional) {
            if (m <= (UINT64_MAX - (uint64_t)d) / 10) {
                m = m * 10 + (uint64_t)d;
                if (fractional) --e;
            } else if (!fractional) {
                ++e;
            }
        };
        while (i < s.size() && s[i] >= '0' && s[i] <= '9') push_digit(s[i++] - '0', false);
        if (i < s.size() && s[i] == '.') {
            ++i;
            if (i >= s.size() || s[i] < '0' || s[i] > '9') return fail("expected fraction digit");
            while (i < s.size() && s[i] >= '0' && s[i] <= '9') push_digit(s[i++] - '0', true);
        }
        if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
            ++i;
            bool eneg = false;
            if (i < s.size() && (s[i] == '+' || s[i] == '-')) eneg = s[i++] == '-';
            if (i >= s.size() || s[i] < '0' || s[i] > '9') return fail("expected exponent digit");
            int x = 0;
            while (i < s.size() && s[i] >= '0' && s[i] <= '9') {
                if (x < 100000) x = x * 10 + (s[i] - '0');
                ++i;
            }
            e += eneg ? -x : x;
        }
        out->mantissa = m;
        out->exponent = e;
        return true;
    }

    bool parse_value(JsonValue* out, int depth) {
        if (depth > 512) return fail("nesting too deep");
        skip_ws();
        if (i >= s.size()) return fail("unexpected end of input");
        char c = s[i];
        if (c == '{') {
            ++i;
            out->kind = JsonValue::Object;
            skip_ws();
            if (i < s.size() && s[i] == '}') {
                ++i;
                return true;
            }
            for (;;) {
                skip_ws();
                std::string key;
                if (!parse_string(&key)) return false;
                skip_ws();
                if (i >= s.size() || s[i] != ':') return fail("expected ':'");
                ++i;
                JsonValue v;
                if (!parse_value(&v, depth + 1)) return false;
                out->entries.emplace_back(std::move(key), std::move(v));
                skip_ws();
                if (i < s.size() && s[i] == ',') {
                    ++i;
                    continue;
                }
                if (i < s.size() && s[i] == '}') {
                    ++i;
                    return true;
                }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++i;
            out->kind = JsonValue::Array;
            skip_ws();
            if (i < s.size() && s[i] == ']') {
                ++i;
                return true;
            }
            for (;;) {
                JsonValue v;
                if (!parse_value(&v, depth + 1)) return false;
                out->items.push_back(std::move(v));
                skip_ws();
                if (i < s.size() && s[i] == ',') {
                    ++i;
                    continue;
                }
                if (i < s.size() && s[i] == ']') {
                    ++i;
                    return true;
                }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') {
            out->kind = JsonValue::String;
            return parse_string(&out->str);
        }
        if (c == 't') {
            out->kind = JsonValue::Bool;
            out->boolean = true;
            return literal("true");
        }
        if (c == 'f') {
            out->kind = JsonValue::Bool;
            out->boolean = false;
            return literal("false");
        }
        if (c == 'n') {
            out->kind = JsonValue::Null;
            return literal("null");
        }
        if (c == '-' || (c >= '0' && c <= '9')) return parse_number(out);
        return fail("unexpected character");
    }
};

void dump_into(const JsonValue& v, std::string* out) {
    switch (v.kind) {
    case JsonValue::Null: *out += "null"; break;
    case JsonValue::Bool: *out += v.boolean ? "true" : "false"; break;
    case JsonValue::Number: {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%.17g", v.as_f64());
        *out += buf;
        break;
    }
    case JsonValue::String: *out += '"' + v.str + '"'; break;
    case JsonValue::Array: {
        *out += '[';
        for (size_t k = 0; k < v.items.size(); ++k) {
            if (k) *out += ',';
            dump_into(v.items[k], out);
        }
        *out += ']';
        break;
    }
    case JsonValue::Object: {
        *out += '{';
        for (size_t k = 0; k < v.entries.size(); ++k) {
            if (k) *out += ',';
            *out += '"' + v.entries[k].first + "\":";
            dump_into(v.entries[k].second, out);
        }
        *out += '}';
        break;
    }
    }
}

} // namespace

double JsonValue::as_f64() const {
    // json crate, `impl From<Number> for f64`
    double n = (double)mantissa;
    int e = exponent;
    if (e < -308) {
        n = exponent_to_power(e + 308) * n;
        e = -308;
    }
    double f = n * exponent_to_power(e);
    return negative ? -f : f;
}

bool JsonValue::as_u64(uint64_t* out) const {
    if (kind != Number) return false;
    if (negative && mantissa != 0) return false;
    uint64_t m = mantissa;
    int e = exponent;
    while (e < 0) { // only exact integers
        if (m % 10 != 0) return false;
        m /= 10;
        ++e;
    }
    while (e > 0) {
        if (m > UINT64_MAX / 10) return false;
        m *= 10;
        --e;
    }
    *out = m;
    return true;
}

const JsonValue* JsonValue::get(const std::string& key) const {
    for (const auto& kv : entries)
        if (kv.first == key) return &kv.second;
    return nullptr;
}

std::string JsonValue::dump() const {
    std::string out;
    dump_into(*this, &out);
    if (out.size() > 400) out = out.substr(0, 400) + "...";
    return out;
}

bool json_parse(const std::string& text, JsonValue* out, std::string* error) {
    Reader r(text);
    JsonValue v;
    if (!r.parse_value(&v, 0)) {
        *error = r.err;
        return false;
    }
    r.skip_ws();
    if (r.i != text.size()) {
        r.fail("trailing characters");
        *error = r.err;
        return false;
    }
    *out = std::move(v);
    return true;
}

} // namespace eucl
