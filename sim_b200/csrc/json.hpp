// sim_b200/csrc/json.hpp -- minimal JSON reader for the wire format of WebSimulation::post_json
// (sim/src/simulator/web.rs:23-31; schema SURVEY.md Appendix E).  Objects keep insertion order;
// numbers keep both their f64 value and, when integral, an exact u64.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

namespace simb {
namespace json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    bool is_uint = false;
    uint64_t u = 0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    const Value* get(const std::string& key) const {
        if (kind != Object) return nullptr;
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_object() const { return kind == Object; }
    bool is_array() const { return kind == Array; }
    bool is_string() const { return kind == String; }
    bool is_number() const { return kind == Number; }
};

struct Parser {
    const char* p;
    const char* end;
    std::string err;

    explicit Parser(const char* s) : p(s), end(s + std::char_traits<char>::length(s)) {}

    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool fail(const char* m) {
        if (err.empty()) err = std::string(m) + " at offset " + std::to_string((long)(p - (end - (end - p))));
        return false;
    }
    bool parse(Value& out) {
        ws();
        if (!value(out, 0)) return false;
        ws();
        if (p != end) return fail("trailing characters");
        return true;
    }
    bool value(Value& v, int depth) {
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end");
        char c = *p;
        if (c == '{') return object(v, depth);
        if (c == '[') return array(v, depth);
        if (c == '"') {
            v.kind = Value::String;
            return string(v.str);
        }
        if (c == 't' && end - p >= 4 && std::string(p, 4) == "true") {
            p += 4;
            v.kind = Value::Bool;
            v.b = true;
            return true;
        }
        if (c == 'f' && end - p >= 5 && std::string(p, 5) == "false") {
            p += 5;
            v.kind = Value::Bool;
            v.b = false;
            return true;
        }
        if (c == 'n' && end - p >= 4 && std::string(p, 4) == "null") {
            p += 4;
            v.kind = Value::Null;
            return true;
        }
        return number(v);
    }
    bool number(Value& v) {
        const char* s = p;
        if (p < end && (*p == '-' || *p == '+')) p++;
        bool digits = false, integral = true;
        while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) {
            if (*p >= '0' && *p <= '9') digits = true;
            else integral = false;
            p++;
        }
        if (!digits) return fail("invalid number");
        std::string t(s, p);
        v.kind = Value::Number;
        v.num = std::strtod(t.c_str(), nullptr);
        if (integral && t[0] != '-') {
            v.is_uint = true;
            v.u = std::strtoull(t.c_str(), nullptr, 10);
        }
        return true;
    }
    bool string(std::string& out) {
        p++;  // opening quote
        out.clear();
        while (p < end && *p != '"') {
            if (*p == '\\') {
                p++;
                if (p >= end) return fail("bad escape");
                switch (*p) {
                    case 'n': out.push_back('\n'); break;
                    case 't': out.push_back('\t'); break;
                    case 'r': out.push_back('\r'); break;
                    case 'b': out.push_back('\b'); break;
                    case 'f': out.push_back('\f'); break;
                    case 'u': {
                        if (end - p < 5) return fail("bad \\u escape");
                        unsigned cp = (unsigned)std::strtoul(std::string(p + 1, 4).c_str(), nullptr, 16);
                        p += 4;
                        if (cp < 0x80) out.push_back((char)cp);
                        else if (cp < 0x800) {
                            out.push_back((char)(0xC0 | (cp >> 6)));
                            out.push_back((char)(0x80 | (cp & 0x3F)));
                        } else {
                            out.push_back((char)(0xE0 | (cp >> 12)));
                            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
                            out.push_back((char)(0x80 | (cp & 0x3F)));
                        }
                        break;
                    }
                    default: out.push_back(*p);
                }
                p++;
            } else {
                out.push_back(*p++);
            }
        }
        if (p >= end) return fail("unterminated string");
        p++;
        return true;
    }
    bool array(Value& v, int depth) {
        v.kind = Value::Array;
        p++;
        ws();
        if (p < end && *p == ']') {
            p++;
            return true;
        }
        for (;;) {
            Value e;
            if (!value(e, depth + 1)) return false;
            v.arr.push_back(std::move(e));
            ws();
            if (p < end && *p == ',') {
                p++;
                continue;
            }
            if (p < end && *p == ']') {
                p++;
                return true;
            }
            return fail("expected , or ]");
        }
    }
    bool object(Value& v, int depth) {
        v.kind = Value::Object;
        p++;
        ws();
        if (p < end && *p == '}') {
            p++;
            return true;
        }
        for (;;) {
            ws();
            if (p >= end || *p != '"') return fail("expected object key");
            std::string k;
            if (!string(k)) return false;
            ws();
            if (p >= end || *p != ':') return fail("expected :");
            p++;
            Value e;
            if (!value(e, depth + 1)) return false;
            v.obj.emplace_back(std::move(k), std::move(e));
            ws();
            if (p < end && *p == ',') {
                p++;
                continue;
            }
            if (p < end && *p == '}') {
                p++;
                return true;
            }
            return fail("expected , or }");
        }
    }
};

}  // namespace json
}  // namespace simb
