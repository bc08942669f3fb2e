// json.hpp -- the little JSON this host layer needs: vectors.meta.json (serde_json's rendering of a
// HashMap<usize, String>: one flat object with stringified integer keys, reference local.rs:155-161) and the header
// of a .safetensors file.  Parser: objects, arrays, strings (with escapes), numbers, true/false/null.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace memex {
namespace json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;   // insertion order kept

    const Value *get(const std::string &key) const
    {
        for (const auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

class Parser {
public:
    explicit Parser(const std::string &s) : s_(s) {}
    Value parse()
    {
        Value v = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string &s_;
    size_t p_ = 0;
    [[noreturn]] void fail(const char *what) const { throw std::runtime_error(std::string("json: ") + what + " at offset " + std::to_string(p_)); }
    void ws()
    {
        while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) ++p_;
    }
    static void utf8(uint32_t cp, std::string &out)
    {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    }
    uint32_t hex4()
    {
        if (p_ + 4 > s_.size()) fail("short \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = s_[p_++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string string()
    {
        if (s_[p_] != '"') fail("expected string");
        ++p_;
        std::string out;
        while (true) {
            if (p_ >= s_.size()) fail("unterminated string");
            char c = s_[p_++];
            if (c == '"') break;
            if (c != '\\') { out += c; continue; }
            if (p_ >= s_.size()) fail("unterminated escape");
            char e = s_[p_++];
            switch (e) {
                case '"': out += '"'; break;
                case '\\': out += '\\'; break;
                case '/': out += '/'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'n': out += '\n'; break;
                case 'r': out += '\r'; break;
                case 't': out += '\t'; break;
                case 'u': {
                    uint32_t cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && p_ + 1 < s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') {
                        p_ += 2;
                        uint32_t lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    utf8(cp, out);
                    break;
                }
                default: fail("bad escape");
            }
        }
        return out;
    }
    Value value()
    {
        ws();
        if (p_ >= s_.size()) fail("unexpected end");
        Value v;
        char c = s_[p_];
        if (c == '{') {
            v.kind = Value::Object;
            ++p_;
            ws();
            if (p_ < s_.size() && s_[p_] == '}') { ++p_; return v; }
            while (true) {
                ws();
                std::string k = string();
                ws();
                if (p_ >= s_.size() || s_[p_] != ':') fail("expected ':'");
                ++p_;
                v.obj.emplace_back(std::move(k), value());
                ws();
                if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
                if (p_ < s_.size() && s_[p_] == '}') { ++p_; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = Value::Array;
            ++p_;
            ws();
            if (p_ < s_.size() && s_[p_] == ']') { ++p_; return v; }
            while (true) {
                v.arr.push_back(value());
                ws();
                if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
                if (p_ < s_.size() && s_[p_] == ']') { ++p_; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.kind = Value::String;
            v.str = string();
        } else if (s_.compare(p_, 4, "true") == 0) { v.kind = Value::Bool; v.b = true; p_ += 4; }
        else if (s_.compare(p_, 5, "false") == 0) { v.kind = Value::Bool; p_ += 5; }
        else if (s_.compare(p_, 4, "null") == 0) { p_ += 4; }
        else {
            char *end = nullptr;
            v.num = std::strtod(s_.c_str() + p_, &end);
            if (end == s_.c_str() + p_) fail("unexpected character");
            v.kind = Value::Number;
            p_ = (size_t)(end - s_.c_str());
        }
        return v;
    }
};

inline Value parse(const std::string &s) { return Parser(s).parse(); }

// serde_json's string escaping: \" \\ \b \f \n \r \t and \u00XX for the other control characters
inline void escape_into(const std::string &s, std::string &out)
{
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
                    static const char *hex = "0123456789abcdef";
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
}  // namespace memex
