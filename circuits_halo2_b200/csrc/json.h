// Minimal JSON reader for the constraint-system description handed to sb_pk_create
// (objects, arrays, strings without escapes beyond \" and \\, integers, true/false/null).
#pragma once
#include <stdlib.h>

#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace sb {
namespace json {

struct Value;
typedef std::shared_ptr<Value> ValueP;
struct Value {
    enum Type { NUL, BOOL, NUM, STR, ARR, OBJ } type = NUL;
    bool b = false;
    long long num = 0;
    std::string str;
    std::vector<ValueP> arr;
    std::map<std::string, ValueP> obj;
    const Value &at(const std::string &k) const {
        auto it = obj.find(k);
        if (type != OBJ || it == obj.end()) throw std::runtime_error("json: missing key '" + k + "'");
        return *it->second;
    }
    bool has(const std::string &k) const { return type == OBJ && obj.count(k); }
    const Value &operator[](size_t i) const {
        if (type != ARR || i >= arr.size()) throw std::runtime_error("json: index out of range");
        return *arr[i];
    }
    size_t size() const { return arr.size(); }
    long long as_int() const {
        if (type != NUM) throw std::runtime_error("json: number expected");
        return num;
    }
    const std::string &as_str() const {
        if (type != STR) throw std::runtime_error("json: string expected");
        return str;
    }
};

class Parser {
   public:
    explicit Parser(const std::string &s) : s_(s), i_(0) {}
    ValueP parse() {
        ValueP v = value();
        ws();
        if (i_ != s_.size()) throw std::runtime_error("json: trailing characters");
        return v;
    }

   private:
    const std::string &s_;
    size_t i_;
    void ws() {
        while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) i_++;
    }
    char peek() {
        ws();
        if (i_ >= s_.size()) throw std::runtime_error("json: unexpected end");
        return s_[i_];
    }
    void expect(char c) {
        if (peek() != c) throw std::runtime_error(std::string("json: expected '") + c + "'");
        i_++;
    }
    ValueP value() {
        char c = peek();
        auto v = std::make_shared<Value>();
        if (c == '{') {
            v->type = Value::OBJ;
            i_++;
            if (peek() == '}') { i_++; return v; }
            while (true) {
                ValueP k = value();
                if (k->type != Value::STR) throw std::runtime_error("json: object key must be a string");
                expect(':');
                v->obj[k->str] = value();
                if (peek() == ',') { i_++; continue; }
                expect('}');
                break;
            }
        } else if (c == '[') {
            v->type = Value::ARR;
            i_++;
            if (peek() == ']') { i_++; return v; }
            while (true) {
                v->arr.push_back(value());
                if (peek() == ',') { i_++; continue; }
                expect(']');
                break;
            }
        } else if (c == '"') {
            v->type = Value::STR;
            i_++;
            while (i_ < s_.size() && s_[i_] != '"') {
                if (s_[i_] == '\\' && i_ + 1 < s_.size()) i_++;
                v->str.push_back(s_[i_++]);
            }
            if (i_ >= s_.size()) throw std::runtime_error("json: unterminated string");
            i_++;
        } else if (c == 't' && s_.compare(i_, 4, "true") == 0) {
            v->type = Value::BOOL; v->b = true; i_ += 4;
        } else if (c == 'f' && s_.compare(i_, 5, "false") == 0) {
            v->type = Value::BOOL; v->b = false; i_ += 5;
        } else if (c == 'n' && s_.compare(i_, 4, "null") == 0) {
            i_ += 4;
        } else {
            size_t j = i_;
            if (s_[j] == '-') j++;
            while (j < s_.size() && s_[j] >= '0' && s_[j] <= '9') j++;
            if (j == i_) throw std::runtime_error("json: unexpected character");
            v->type = Value::NUM;
            v->num = atoll(s_.substr(i_, j - i_).c_str());
            i_ = j;
        }
        return v;
    }
};

inline ValueP parse(const std::string &s) { return Parser(s).parse(); }

}  // namespace json
}  // namespace sb
