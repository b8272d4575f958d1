// Minimal JSON reader for the scene front end.
//
// Numbers follow the conversion of the `json` crate the reference uses (json 0.11.13, pinned in
// Cargo.lock; source not vendored -> RECOLLECTION, see oracle/ASSUMPTIONS.md): the literal is kept
// as (sign, u64 mantissa, i16 decimal exponent) and converted with ONE multiply by a power-of-ten
// table entry, which is not always the correctly rounded strtod value (e.g. `1.458`).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace eucl {

struct JsonValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool boolean = false;
    // number parts (json crate: Number::from_parts)
    bool negative = false;
    uint64_t mantissa = 0;
    int exponent = 0;
    std::string str;
    std::vector<JsonValue> items;                           // Array
    std::vector<std::pair<std::string, JsonValue>> entries; // Object, insertion order

    double as_f64() const;           // kind must be Number
    bool as_u64(uint64_t* out) const; // non-negative integer that fits
    const JsonValue* get(const std::string& key) const;
    std::string dump() const; // compact, for error messages
};

// Returns false and sets *error on a syntax error.
bool json_parse(const std::string& text, JsonValue* out, std::string* error);

} // namespace eucl
