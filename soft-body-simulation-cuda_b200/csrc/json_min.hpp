// Minimal JSON reader for context.json (the reference parses it with nlohmann::json,
// src/context.cpp:319-387).  Supports objects, arrays, strings, numbers, true/false/null.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace pdb200 {

struct Json {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;   // insertion order kept

    bool contains(const std::string& k) const { return find(k) != nullptr; }
    const Json* find(const std::string& k) const {
        if (type != Obj) return nullptr;
        for (auto& kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const Json& at(const std::string& k) const {
        const Json* j = find(k);
        if (!j) throw std::runtime_error("json: missing key '" + k + "'");
        return *j;
    }
    const Json& operator[](size_t i) const {
        if (type != Arr || i >= arr.size()) throw std::runtime_error("json: bad array index");
        return arr[i];
    }
    size_t size() const { return type == Arr ? arr.size() : type == Obj ? obj.size() : 0; }
    bool is_null() const { return type == Null; }
    double number() const {
        if (type != Num) throw std::runtime_error("json: not a number");
        return num;
    }
    const std::string& string() const {
        if (type != Str) throw std::runtime_error("json: not a string");
        return str;
    }
    bool boolean() const {
        if (type != Bool) throw std::runtime_error("json: not a bool");
        return b;
    }
    // nlohmann-like value(key, default)
    double value(const std::string& k, double d) const { auto j = find(k); return (j && j->type == Num) ? j->num : d; }
    bool value(const std::string& k, bool d) const { auto j = find(k); return (j && j->type == Bool) ? j->b : d; }
    std::string value(const std::string& k, const char* d) const { auto j = find(k); return (j && j->type == Str) ? j->str : std::string(d); }
};

class JsonParser {
public:
    explicit JsonParser(const std::string& text) : s(text) {}
    Json parse() {
        Json j = value();
        ws();
        if (p != s.size()) fail("trailing characters");
        return j;
    }

private:
    const std::string& s;
    size_t p = 0;
    [[noreturn]] void fail(const char* m) const {
        throw std::runtime_error(std::string("json parse error at byte ") + std::to_string(p) + ": " + m);
    }
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) p++; }
    bool lit(const char* w) {
        size_t n = 0; while (w[n]) n++;
        if (s.compare(p, n, w) == 0) { p += n; return true; }
        return false;
    }
    Json value() {
        ws();
        if (p >= s.size()) fail("unexpected end");
        char c = s[p];
        Json j;
        if (c == '{') {
            j.type = Json::Obj; p++; ws();
            if (p < s.size() && s[p] == '}') { p++; return j; }
            for (;;) {
                ws();
                if (p >= s.size() || s[p] != '"') fail("expected key");
                std::string k = str();
                ws();
                if (p >= s.size() || s[p] != ':') fail("expected ':'");
                p++;
                j.obj.emplace_back(std::move(k), value());
                ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == '}') { p++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            j.type = Json::Arr; p++; ws();
            if (p < s.size() && s[p] == ']') { p++; return j; }
            for (;;) {
                j.arr.push_back(value());
                ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == ']') { p++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            j.type = Json::Str; j.str = str();
        } else if (lit("true")) { j.type = Json::Bool; j.b = true;
        } else if (lit("false")) { j.type = Json::Bool; j.b = false;
        } else if (lit("null")) { j.type = Json::Null;
        } else {
            const char* b = s.c_str() + p; char* e = nullptr;
            double v = std::strtod(b, &e);
            if (e == b) fail("bad token");
            p += (size_t)(e - b);
            j.type = Json::Num; j.num = v;
        }
        return j;
    }
    std::string str() {
        std::string out; p++;  // opening quote
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\' && p + 1 < s.size()) {
                char e = s[p + 1]; p += 2;
                switch (e) {
                    case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': { if (p + 4 > s.size()) fail("bad \\u"); unsigned cp = (unsigned)std::strtoul(s.substr(p, 4).c_str(), nullptr, 16); p += 4;
                        if (cp < 0x80) out += (char)cp; else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); } break; }
                    default: out += e;
                }
            } else out += s[p++];
        }
        if (p >= s.size()) fail("unterminated string");
        p++;
        return out;
    }
};

}  // namespace pdb200
