// sim_b200/csrc/json_write.hpp -- JSON text for the string API of WebSimulation
// (sim/src/simulator/web.rs:87-215): serde_json::to_string of Vec<Message> (coupling.rs:64-71,
// camelCase keys, field order as declared) and Vec<ModelRecord> (models/mod.rs:48-52).
// Numbers are written the way serde_json writes an f64: the shortest digits that round-trip, laid
// out by the rules of the `ryu` crate it uses (always a fraction or an exponent; `null` when not
// finite).
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <string>

namespace simb {
namespace json {

inline void write_f64(std::string& out, double v) {
    if (!std::isfinite(v)) {
        out += "null";
        return;
    }
    if (std::signbit(v)) {
        out += '-';
        v = -v;
    }
    if (v == 0.0) {
        out += "0.0";
        return;
    }
    // shortest round-trip digits in scientific form: d[.ddd]e[+-]xx
    char buf[40];
    auto res = std::to_chars(buf, buf + sizeof buf - 1, v, std::chars_format::scientific);
    *res.ptr = '\0';
    std::string digits;
    int exp10 = 0;
    {
        const char* p = buf;
        for (; p < res.ptr && *p != 'e'; p++)
            if (*p != '.') digits += *p;
        if (p < res.ptr) exp10 = (int)std::strtol(p + 1, nullptr, 10);
    }
    const int len = (int)digits.size();
    const int kk = exp10 + 1;       // position of the decimal point relative to the first digit
    const int k = kk - len;         // value = digits * 10^k
    if (k >= 0 && kk <= 16) {       // 1234e7 -> 12340000000.0
        out += digits;
        out.append((size_t)k, '0');
        out += ".0";
    } else if (kk > 0 && kk <= 16) {  // 1234e-2 -> 12.34
        out.append(digits, 0, (size_t)kk);
        out += '.';
        out.append(digits, (size_t)kk, std::string::npos);
    } else if (kk > -5 && kk <= 0) {  // 1234e-6 -> 0.001234
        out += "0.";
        out.append((size_t)(-kk), '0');
        out += digits;
    } else {                          // 1e30, 1.234e33, 1.5e-7
        out += digits[0];
        if (len > 1) {
            out += '.';
            out.append(digits, 1, std::string::npos);
        }
        out += 'e';
        out += std::to_string(kk - 1);
    }
}

inline void write_string(std::string& out, const std::string& s) {
    static const char* hex = "0123456789abcdef";
    out += '"';
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) {
                    out += "\\u00";
                    out += hex[c >> 4];
                    out += hex[c & 15];
                } else {
                    out += (char)c;
                }
        }
    }
    out += '"';
}

}  // namespace json
}  // namespace simb
