// sim_b200/csrc/yaml.hpp -- the YAML side of WebSimulation's string API (sim/src/simulator/web.rs:56-81, 91-215:
// post_yaml / put_yaml / get_yaml and the *_yaml accessors), as a converter between YAML text and the JSON value
// tree of json.hpp.
//
// Reader: the block-style subset that serde_yaml 0.8 emits and that the reference's fixtures use
// (sim/tests/web.rs:322-407): block mappings and sequences by indentation, `- key: value` compact entries, plain /
// single- / double-quoted scalars, flow collections `[a, b]` / `{a: b}` (nested, over several lines too), `~` / null, booleans,
// integers, floats, `.inf` / `-.inf` / `.nan`, `#` comments, a leading `---`.  Anchors, tags, block scalars and
// multi-document streams are not part of that wire format and are rejected.
// Writer: the layout of serde_yaml 0.8 (`---` header, two-space indentation, sequences indented under their key,
// `~` for null, `.inf`, strings quoted only when a plain scalar would be read back as something else).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "json.hpp"
#include "json_write.hpp"

namespace simb {
namespace yaml {

using json::Value;

struct Line {
    int indent;
    std::string text;  // without indentation, comments and trailing blanks
    int number;
};

inline bool is_plain_number(const std::string& s, Value* out) {
    if (s.empty()) return false;
    size_t i = 0;
    if (s[i] == '-' || s[i] == '+') i++;
    if (i >= s.size()) return false;
    bool digits = false, frac = false, exp = false;
    const size_t start = i;
    for (; i < s.size(); i++) {
        const char c = s[i];
        if (c >= '0' && c <= '9') digits = true;
        else if (c == '.' && !frac && !exp) frac = true;
        else if ((c == 'e' || c == 'E') && digits && !exp) {
            exp = true;
            if (i + 1 < s.size() && (s[i + 1] == '-' || s[i + 1] == '+')) i++;
        } else return false;
    }
    if (!digits) return false;
    (void)start;
    out->kind = Value::Number;
    out->num = std::strtod(s.c_str(), nullptr);
    out->is_uint = !frac && !exp && s[0] != '-';
    if (out->is_uint) out->u = std::strtoull(s[0] == '+' ? s.c_str() + 1 : s.c_str(), nullptr, 10);
    return true;
}

inline bool unquote(const std::string& s, std::string* out, std::string* err) {
    out->clear();
    const char q = s[0];
    for (size_t i = 1; i < s.size(); i++) {
        const char c = s[i];
        if (c == q) {
            if (q == '\'' && i + 1 < s.size() && s[i + 1] == '\'') {
                *out += '\'';
                i++;
                continue;
            }
            if (i + 1 != s.size()) {
                *err = "characters after a closing quote";
                return false;
            }
            return true;
        }
        if (q == '"' && c == '\\' && i + 1 < s.size()) {
            const char e = s[++i];
            switch (e) {
                case 'n': *out += '\n'; break;
                case 't': *out += '\t'; break;
                case 'r': *out += '\r'; break;
                case '0': *out += '\0'; break;
                case '"': *out += '"'; break;
                case '\\': *out += '\\'; break;
                case '/': *out += '/'; break;
                default: *out += e; break;
            }
            continue;
        }
        *out += c;
    }
    *err = "unterminated quoted scalar";
    return false;
}

struct Reader {
    std::vector<Line> lines;
    size_t pos = 0;
    std::string err;

    bool fail(const std::string& m, int line) {
        if (err.empty()) err = "yaml line " + std::to_string(line) + ": " + m;
        return false;
    }
    // strip a trailing comment that is outside quotes
    static std::string strip_comment(const std::string& s) {
        char q = 0;
        for (size_t i = 0; i < s.size(); i++) {
            const char c = s[i];
            if (q) {
                if (c == '\\' && q == '"') i++;
                else if (c == q) q = 0;
            } else if (c == '"' || c == '\'') {
                q = c;
            } else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) {
                return s.substr(0, i);
            }
        }
        return s;
    }
    bool split(const char* text) {
        int number = 0;
        const char* p = text;
        while (*p) {
            const char* e = std::strchr(p, '\n');
            std::string raw = e ? std::string(p, e) : std::string(p);
            p = e ? e + 1 : p + raw.size();
            number++;
            if (!raw.empty() && raw.back() == '\r') raw.pop_back();
            std::string s = strip_comment(raw);
            while (!s.empty() && (s.back() == ' ' || s.back() == '\t')) s.pop_back();
            size_t ind = 0;
            while (ind < s.size() && s[ind] == ' ') ind++;
            if (ind < s.size() && s[ind] == '\t') return fail("tab used for indentation", number);
            if (ind == s.size()) continue;
            std::string body = s.substr(ind);
            if (body == "---" && lines.empty()) continue;
            if (body == "---" || body == "...") return fail("multi-document streams are not supported", number);
            if (open_depth > 0) {  // continuation of a flow collection that did not close on the line it opened
                lines.back().text += ' ';
                lines.back().text += body;
                open_depth += flow_depth(body, 0);
                continue;
            }
            lines.push_back(Line{(int)ind, body, number});
            const size_t v = value_start(body);
            if (v < body.size() && (body[v] == '[' || body[v] == '{')) open_depth = flow_depth(body, v);
        }
        if (open_depth > 0) return fail("unterminated flow collection", number);
        return true;
    }
    int open_depth = 0;
    // where the value of a line begins: after any "- " entries and a "key: " prefix
    static size_t value_start(const std::string& b) {
        size_t i = 0;
        while (i + 1 < b.size() && b[i] == '-' && b[i + 1] == ' ') {
            i += 2;
            while (i < b.size() && b[i] == ' ') i++;
        }
        if (i < b.size() && (b[i] == '[' || b[i] == '{')) return i;
        char q = 0;
        for (size_t k = i; k < b.size(); k++) {
            const char c = b[k];
            if (q) {
                if (c == '\\' && q == '"') k++;
                else if (c == q) q = 0;
            } else if ((c == '"' || c == '\'') && k == i) {
                q = c;
            } else if (c == ':' && k + 1 < b.size() && b[k + 1] == ' ') {
                k += 2;
                while (k < b.size() && b[k] == ' ') k++;
                return k;
            }
        }
        return b.size();
    }
    // brackets opened minus brackets closed from position `from`, outside quotes
    static int flow_depth(const std::string& b, size_t from) {
        int d = 0;
        char q = 0;
        for (size_t k = from; k < b.size(); k++) {
            const char c = b[k];
            if (q) {
                if (c == '\\' && q == '"') k++;
                else if (c == q) q = 0;
            } else if (c == '"' || c == '\'') {
                q = c;
            } else if (c == '[' || c == '{') {
                d++;
            } else if (c == ']' || c == '}') {
                d--;
            }
        }
        return d;
    }

    // ---- scalars and one-line flow collections ----
    bool scalar(const std::string& s, Value* out, int line) {
        *out = Value();
        if (s.empty() || s == "~" || s == "null" || s == "Null" || s == "NULL") return true;
        if (s[0] == '"' || s[0] == '\'') {
            out->kind = Value::String;
            std::string e;
            if (!unquote(s, &out->str, &e)) return fail(e, line);
            return true;
        }
        if (s[0] == '[' || s[0] == '{') return flow(s, out, line);
        if (s[0] == '&' || s[0] == '*' || s[0] == '!' || s[0] == '|' || s[0] == '>')
            return fail("anchors, tags and block scalars are not supported", line);
        if (s == "true" || s == "True" || s == "TRUE") {
            out->kind = Value::Bool;
            out->b = true;
            return true;
        }
        if (s == "false" || s == "False" || s == "FALSE") {
            out->kind = Value::Bool;
            out->b = false;
            return true;
        }
        if (s == ".inf" || s == ".Inf" || s == ".INF" || s == "+.inf") {
            out->kind = Value::Number;
            out->num = HUGE_VAL;
            return true;
        }
        if (s == "-.inf" || s == "-.Inf" || s == "-.INF") {
            out->kind = Value::Number;
            out->num = -HUGE_VAL;
            return true;
        }
        if (s == ".nan" || s == ".NaN" || s == ".NAN") {
            out->kind = Value::Number;
            out->num = std::nan("");
            return true;
        }
        if (is_plain_number(s, out)) return true;
        out->kind = Value::String;
        out->str = s;
        return true;
    }
    // split the inside of a flow collection at top-level commas
    static std::vector<std::string> flow_items(const std::string& inner) {
        std::vector<std::string> items;
        std::string cur;
        int depth = 0;
        char q = 0;
        for (size_t i = 0; i < inner.size(); i++) {
            const char c = inner[i];
            if (q) {
                cur += c;
                if (c == '\\' && q == '"' && i + 1 < inner.size()) cur += inner[++i];
                else if (c == q) q = 0;
                continue;
            }
            if (c == '"' || c == '\'') q = c;
            if (c == '[' || c == '{') depth++;
            if (c == ']' || c == '}') depth--;
            if (c == ',' && depth == 0) {
                items.push_back(cur);
                cur.clear();
                continue;
            }
            cur += c;
        }
        items.push_back(cur);
        for (auto& it : items) {
            size_t a = 0, b = it.size();
            while (a < b && it[a] == ' ') a++;
            while (b > a && it[b - 1] == ' ') b--;
            it = it.substr(a, b - a);
        }
        if (items.size() == 1 && items[0].empty()) items.clear();
        return items;
    }
    bool flow(const std::string& s, Value* out, int line) {
        const char open = s[0], close = open == '[' ? ']' : '}';
        if (s.back() != close) return fail("text after the end of a flow collection", line);
        const std::vector<std::string> items = flow_items(s.substr(1, s.size() - 2));
        if (open == '[') {
            out->kind = Value::Array;
            for (auto& it : items) {
                Value v;
                if (!scalar(it, &v, line)) return false;
                out->arr.push_back(v);
            }
            return true;
        }
        out->kind = Value::Object;
        for (auto& it : items) {
            std::string key, rest;
            if (!key_of(it, &key, &rest, line)) return fail("flow mapping entries need `key: value`", line);
            Value v;
            if (!scalar(rest, &v, line)) return false;
            out->obj.push_back({key, v});
        }
        return true;
    }
    // "key: rest" / "key:" (key plain or quoted); false when the text is not a mapping entry
    bool key_of(const std::string& s, std::string* key, std::string* rest, int line) {
        size_t i = 0;
        if (s[0] == '"' || s[0] == '\'') {
            const char q = s[0];
            for (i = 1; i < s.size(); i++) {
                if (s[i] == '\\' && q == '"') i++;
                else if (s[i] == q) break;
            }
            if (i >= s.size()) return false;
            std::string e;
            if (!unquote(s.substr(0, i + 1), key, &e)) return fail(e, line);
            i++;
            if (i >= s.size() || s[i] != ':') return false;
        } else {
            if (s[0] == '[' || s[0] == '{') return false;
            for (;; i++) {
                if (i >= s.size()) return false;
                if (s[i] == ':' && (i + 1 == s.size() || s[i + 1] == ' ')) break;
            }
            *key = s.substr(0, i);
        }
        size_t r = i + 1;
        while (r < s.size() && s[r] == ' ') r++;
        *rest = s.substr(r);
        return true;
    }

    // ---- block structure ----
    bool node(int indent, Value* out) {  // the node starting at lines[pos], which is indented by `indent`
        const Line& ln = lines[pos];
        if (ln.text[0] == '-' && (ln.text.size() == 1 || ln.text[1] == ' ')) return sequence(indent, out);
        std::string key, rest;
        if (key_of(ln.text, &key, &rest, ln.number)) return mapping(indent, out);
        if (!err.empty()) return false;
        pos++;
        return scalar(ln.text, out, ln.number);
    }
    bool value_after(const std::string& rest, int parent_indent, Value* out, int line, bool seq_allowed_at_parent) {
        if (!rest.empty()) return scalar(rest, out, line);
        if (pos < lines.size()) {  // nested block (a sequence may sit at the parent's own indentation)
            const Line& nx = lines[pos];
            const bool dash = nx.text[0] == '-' && (nx.text.size() == 1 || nx.text[1] == ' ');
            if (nx.indent > parent_indent || (seq_allowed_at_parent && dash && nx.indent == parent_indent)) return node(nx.indent, out);
        }
        *out = Value();  // `key:` with nothing under it = null
        return true;
    }
    bool mapping(int indent, Value* out) {
        out->kind = Value::Object;
        while (pos < lines.size() && lines[pos].indent == indent) {
            const Line ln = lines[pos];
            if (ln.text[0] == '-' && (ln.text.size() == 1 || ln.text[1] == ' ')) break;
            std::string key, rest;
            if (!key_of(ln.text, &key, &rest, ln.number)) return fail("expected `key: value`", ln.number);
            pos++;
            Value v;
            if (!value_after(rest, indent, &v, ln.number, true)) return false;
            out->obj.push_back({key, v});
        }
        if (pos < lines.size() && lines[pos].indent > indent) return fail("unexpected indentation", lines[pos].number);
        return true;
    }
    bool sequence(int indent, Value* out) {
        out->kind = Value::Array;
        while (pos < lines.size() && lines[pos].indent == indent) {
            Line& ln = lines[pos];
            if (!(ln.text[0] == '-' && (ln.text.size() == 1 || ln.text[1] == ' '))) break;
            size_t r = 1;
            while (r < ln.text.size() && ln.text[r] == ' ') r++;
            const std::string rest = ln.text.substr(r);
            Value v;
            if (rest.empty()) {
                pos++;
                if (!value_after("", indent, &v, ln.number, false)) return false;
            } else {
                // "- key: value" opens a mapping whose entries are indented to the column of `key`; "- - x" a sequence
                ln.indent = indent + (int)r;
                ln.text = rest;
                if (!node(ln.indent, &v)) return false;
            }
            out->arr.push_back(v);
        }
        return true;
    }
    bool parse(const char* text, Value* out) {
        if (!split(text)) return false;
        if (lines.empty()) {
            *out = Value();
            return true;
        }
        if (!node(lines[0].indent, out)) return false;
        if (pos != lines.size()) return fail("unexpected content", lines[pos].number);
        return true;
    }
};

inline bool to_value(const char* text, Value* out, std::string* err) {
    Reader r;
    if (r.parse(text, out)) return true;
    if (err) *err = r.err;
    return false;
}

// ---- writers -----------------------------------------------------------------------------------------------
inline void write_json(std::string& o, const Value& v, int indent = -1, int depth = 0) {  // indent < 0: compact
    auto nl = [&](int d) {
        if (indent < 0) return;
        o += '\n';
        o.append((size_t)(d * indent), ' ');
    };
    switch (v.kind) {
        case Value::Null: o += "null"; break;
        case Value::Bool: o += v.b ? "true" : "false"; break;
        case Value::Number:
            if (v.is_uint) o += std::to_string(v.u);
            else json::write_f64(o, v.num);
            break;
        case Value::String: json::write_string(o, v.str); break;
        case Value::Array:
            if (v.arr.empty()) {
                o += "[]";
                break;
            }
            o += '[';
            for (size_t i = 0; i < v.arr.size(); i++) {
                if (i) o += ',';
                nl(depth + 1);
                write_json(o, v.arr[i], indent, depth + 1);
            }
            nl(depth);
            o += ']';
            break;
        case Value::Object:
            if (v.obj.empty()) {
                o += "{}";
                break;
            }
            o += '{';
            for (size_t i = 0; i < v.obj.size(); i++) {
                if (i) o += ',';
                nl(depth + 1);
                json::write_string(o, v.obj[i].first);
                o += indent < 0 ? ":" : ": ";
                write_json(o, v.obj[i].second, indent, depth + 1);
            }
            nl(depth);
            o += '}';
            break;
    }
}

inline bool plain_is_safe(const std::string& s) {
    if (s.empty()) return false;
    Value probe;
    if (is_plain_number(s, &probe)) return false;
    static const char* words[] = {"~", "null", "Null", "NULL", "true", "True", "TRUE", "false", "False", "FALSE", "yes", "no", "on",
                                  "off", "Yes", "No", "On", "Off", "y", "n", "Y", "N", ".inf", "-.inf", ".nan", "---", "..."};
    for (const char* w : words)
        if (s == w) return false;
    const char f = s[0];
    if (std::strchr("-?:,[]{}#&*!|>'\"%@` ", f)) return false;
    if (s.back() == ' ' || s.back() == ':') return false;
    for (size_t i = 0; i < s.size(); i++) {
        const unsigned char c = (unsigned char)s[i];
        if (c < 0x20 || c == '"' || c == '\\') return false;
        if (c == ':' && i + 1 < s.size() && s[i + 1] == ' ') return false;
        if (c == '#' && i > 0 && s[i - 1] == ' ') return false;
    }
    return true;
}
inline void write_scalar(std::string& o, const Value& v) {
    switch (v.kind) {
        case Value::Null: o += "~"; break;
        case Value::Bool: o += v.b ? "true" : "false"; break;
        case Value::Number:
            if (v.is_uint) o += std::to_string(v.u);
            else if (std::isnan(v.num)) o += ".nan";
            else if (std::isinf(v.num)) o += v.num > 0 ? ".inf" : "-.inf";
            else json::write_f64(o, v.num);
            break;
        case Value::String:
            if (plain_is_safe(v.str)) o += v.str;
            else json::write_string(o, v.str);
            break;
        default: break;
    }
}
inline void write_yaml_node(std::string& o, const Value& v, int indent, bool inline_first);
inline void write_yaml_entry_value(std::string& o, const Value& v, int indent) {
    if (v.kind == Value::Array && !v.arr.empty()) {
        o += '\n';
        write_yaml_node(o, v, indent + 2, false);
    } else if (v.kind == Value::Object && !v.obj.empty()) {
        o += '\n';
        write_yaml_node(o, v, indent + 2, false);
    } else {
        o += ' ';
        if (v.kind == Value::Array) o += "[]";
        else if (v.kind == Value::Object) o += "{}";
        else write_scalar(o, v);
        o += '\n';
    }
}
// writes the node with every line indented by `indent`; inline_first: the first line continues a "- " prefix
inline void write_yaml_node(std::string& o, const Value& v, int indent, bool inline_first) {
    auto pad = [&](bool first) {
        if (!(first && inline_first)) o.append((size_t)indent, ' ');
    };
    if (v.kind == Value::Array && !v.arr.empty()) {
        for (size_t i = 0; i < v.arr.size(); i++) {
            pad(i == 0);
            o += "- ";
            const Value& e = v.arr[i];
            if ((e.kind == Value::Array && !e.arr.empty()) || (e.kind == Value::Object && !e.obj.empty())) {
                write_yaml_node(o, e, indent + 2, true);
            } else {
                if (e.kind == Value::Array) o += "[]";
                else if (e.kind == Value::Object) o += "{}";
                else write_scalar(o, e);
                o += '\n';
            }
        }
    } else if (v.kind == Value::Object && !v.obj.empty()) {
        for (size_t i = 0; i < v.obj.size(); i++) {
            pad(i == 0);
            Value k;
            k.kind = Value::String;
            k.str = v.obj[i].first;
            write_scalar(o, k);
            o += ':';
            write_yaml_entry_value(o, v.obj[i].second, indent);
        }
    } else {
        pad(true);
        if (v.kind == Value::Array) o += "[]";
        else if (v.kind == Value::Object) o += "{}";
        else write_scalar(o, v);
        o += '\n';
    }
}
inline void write_yaml(std::string& o, const Value& v) {
    o += "---\n";
    write_yaml_node(o, v, 0, false);
}

}  // namespace yaml
}  // namespace simb
