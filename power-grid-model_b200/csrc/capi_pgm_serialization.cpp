// JSON / msgpack (de)serialization of datasets behind the reference's C API names (serialization.h; writable datasets of
// dataset.h): the on-disk / wire format either side of the calculation path.
//   format      docs/advanced_documentation/native-data-interface / serialization: {"version": "1.0", "type": <dataset>,
//               "is_batch": bool, "attributes": {component: [attribute names]}, "data": {component: [rows]} | [scenario maps]};
//               a row is a map attribute -> value or, when the component has predefined attributes, a compact list; NaN / na
//               values are null or absent; +-inf are the strings "inf" / "+inf" / "-inf"; three-phase values are lists of three
//   reading     auxiliary/serialization/deserializer.hpp:600-1169 (required keys, per-component element counts, uniform vs
//               sparse buffers, attribute indications when every row is a compact list, unknown attributes of a map row skipped)
//   writing     auxiliary/serialization/serializer.hpp:281-620 (scenario list with empty components omitted; compact list keeps
//               the attributes that are not NaN over the whole buffer; map rows keep the non-NaN attributes of each element)
// One value tree is the meeting point of both formats: JSON text <-> tree <-> msgpack bytes <-> dataset buffers.
#include "capi_pgm_dataset.hpp"

#include <charconv>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <set>

using namespace pgmb;
using namespace pgmb::capi;

namespace {

struct SerializationError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Value {
    enum Kind { nil, boolean, integer, real, string, array, map } kind{nil};
    bool b{};
    int64_t i{};
    double f{};
    std::string s;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;
    Value const* find(std::string const& key) const {
        for (auto const& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

// ---- JSON text -> tree ------------------------------------------------------------------------------------------------
class JsonReader {
  public:
    JsonReader(char const* data, size_t size) : p_{data}, end_{data + size}, begin_{data} {}
    Value parse_document() {
        Value v = parse_value(0);
        skip_ws();
        if (p_ != end_) fail("Error in parsing");
        return v;
    }

  private:
    char const* p_;
    char const* end_;
    char const* begin_;
    [[noreturn]] void fail(char const* what) const {
        throw SerializationError(std::string(what) + " json at position " + std::to_string(p_ - begin_) + "!\n");
    }
    void skip_ws() {
        while (p_ != end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\r' || *p_ == '\t')) ++p_;
    }
    void expect(char const* lit) {
        size_t const n = std::strlen(lit);
        if (static_cast<size_t>(end_ - p_) < n || std::memcmp(p_, lit, n) != 0) fail("Error in parsing");
        p_ += n;
    }
    Value parse_value(int depth) {
        if (depth > 10) throw SerializationError("Json depth exceeds the limit of 10!\n");
        skip_ws();
        if (p_ == end_) fail("Insufficient bytes in parsing");
        Value v;
        switch (*p_) {
        case '{': {
            ++p_;
            v.kind = Value::map;
            skip_ws();
            if (p_ != end_ && *p_ == '}') {
                ++p_;
                return v;
            }
            while (true) {
                skip_ws();
                if (p_ == end_) fail("Insufficient bytes in parsing");
                if (*p_ != '"') fail("Error in parsing");
                std::string key = parse_string();
                skip_ws();
                if (p_ == end_) fail("Insufficient bytes in parsing");
                if (*p_ != ':') fail("Error in parsing");
                ++p_;
                v.obj.emplace_back(std::move(key), parse_value(depth + 1));
                skip_ws();
                if (p_ == end_) fail("Insufficient bytes in parsing");
                if (*p_ == ',') {
                    ++p_;
                    continue;
                }
                if (*p_ == '}') {
                    ++p_;
                    return v;
                }
                fail("Error in parsing");
            }
        }
        case '[': {
            ++p_;
            v.kind = Value::array;
            skip_ws();
            if (p_ != end_ && *p_ == ']') {
                ++p_;
                return v;
            }
            while (true) {
                v.arr.push_back(parse_value(depth + 1));
                skip_ws();
                if (p_ == end_) fail("Insufficient bytes in parsing");
                if (*p_ == ',') {
                    ++p_;
                    continue;
                }
                if (*p_ == ']') {
                    ++p_;
                    return v;
                }
                fail("Error in parsing");
            }
        }
        case '"':
            v.kind = Value::string;
            v.s = parse_string();
            return v;
        case 't':
            expect("true");
            v.kind = Value::boolean;
            v.b = true;
            return v;
        case 'f':
            expect("false");
            v.kind = Value::boolean;
            v.b = false;
            return v;
        case 'n':
            expect("null");
            return v;
        default: return parse_number();
        }
    }
    Value parse_number() {
        char const* start = p_;
        bool is_real = false;
        if (p_ != end_ && (*p_ == '-' || *p_ == '+')) ++p_;
        while (p_ != end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '-' || *p_ == '+')) {
            is_real = is_real || *p_ == '.' || *p_ == 'e' || *p_ == 'E';
            ++p_;
        }
        if (p_ == start) fail("Error in parsing");
        Value v;
        if (!is_real) {
            int64_t iv = 0;
            auto const r = std::from_chars(start, p_, iv);
            if (r.ec == std::errc{} && r.ptr == p_) {
                v.kind = Value::integer;
                v.i = iv;
                return v;
            }
        }
        double dv = 0.0;
        auto const r = std::from_chars(*start == '+' ? start + 1 : start, p_, dv);
        if (r.ec != std::errc{} || r.ptr != p_) {
            p_ = start;
            fail("Error in parsing");
        }
        v.kind = Value::real;
        v.f = dv;
        return v;
    }
    static void append_utf8(std::string& out, uint32_t cp) {
        if (cp < 0x80) {
            out.push_back(static_cast<char>(cp));
        } else if (cp < 0x800) {
            out.push_back(static_cast<char>(0xC0 | (cp >> 6)));
            out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
        } else if (cp < 0x10000) {
            out.push_back(static_cast<char>(0xE0 | (cp >> 12)));
            out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
        } else {
            out.push_back(static_cast<char>(0xF0 | (cp >> 18)));
            out.push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
        }
    }
    uint32_t hex4() {
        if (end_ - p_ < 4) fail("Insufficient bytes in parsing");
        uint32_t v = 0;
        for (int k = 0; k != 4; ++k, ++p_) {
            char const c = *p_;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= static_cast<uint32_t>(c - '0');
            else if (c >= 'a' && c <= 'f') v |= static_cast<uint32_t>(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= static_cast<uint32_t>(c - 'A' + 10);
            else fail("Error in parsing");
        }
        return v;
    }
    std::string parse_string() {
        ++p_; // opening quote
        std::string out;
        while (true) {
            if (p_ == end_) fail("Insufficient bytes in parsing");
            char const c = *p_++;
            if (c == '"') return out;
            if (c != '\\') {
                out.push_back(c);
                continue;
            }
            if (p_ == end_) fail("Insufficient bytes in parsing");
            char const e = *p_++;
            switch (e) {
            case '"': out.push_back('"'); break;
            case '\\': out.push_back('\\'); break;
            case '/': out.push_back('/'); break;
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            case 'n': out.push_back('\n'); break;
            case 'r': out.push_back('\r'); break;
            case 't': out.push_back('\t'); break;
            case 'u': {
                uint32_t cp = hex4();
                if (cp >= 0xD800 && cp <= 0xDBFF && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                    p_ += 2;
                    uint32_t const lo = hex4();
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                }
                append_utf8(out, cp);
                break;
            }
            default: fail("Error in parsing");
            }
        }
    }
};

// ---- msgpack bytes -> tree ---------------------------------------------------------------------------------------------
class MsgpackReader {
  public:
    MsgpackReader(char const* data, size_t size) : p_{reinterpret_cast<unsigned char const*>(data)}, end_{p_ + size}, begin_{p_} {}
    Value parse_document() { return parse_value(0); }

  private:
    unsigned char const* p_;
    unsigned char const* end_;
    unsigned char const* begin_;
    [[noreturn]] void fail(char const* what) const {
        throw SerializationError(std::string(what) + " msgpack at position " + std::to_string(p_ - begin_) + "!\n");
    }
    void need(size_t n) const {
        if (static_cast<size_t>(end_ - p_) < n) fail("Insufficient bytes in parsing");
    }
    uint64_t be(int n) {
        need(static_cast<size_t>(n));
        uint64_t v = 0;
        for (int k = 0; k != n; ++k) v = (v << 8) | *p_++;
        return v;
    }
    std::string str(size_t n) {
        need(n);
        std::string s(reinterpret_cast<char const*>(p_), n);
        p_ += n;
        return s;
    }
    Value make_array(size_t n, int depth) {
        Value v;
        v.kind = Value::array;
        v.arr.reserve(n);
        for (size_t k = 0; k != n; ++k) v.arr.push_back(parse_value(depth + 1));
        return v;
    }
    Value make_map(size_t n, int depth) {
        Value v;
        v.kind = Value::map;
        v.obj.reserve(n);
        for (size_t k = 0; k != n; ++k) {
            Value key = parse_value(depth + 1);
            if (key.kind != Value::string) fail("Error in parsing (map key is not a string)");
            v.obj.emplace_back(std::move(key.s), parse_value(depth + 1));
        }
        return v;
    }
    Value parse_value(int depth) {
        if (depth > 10) throw SerializationError("Json depth exceeds the limit of 10!\n");
        need(1);
        unsigned const t = *p_++;
        Value v;
        auto integer = [&](int64_t x) {
            v.kind = Value::integer;
            v.i = x;
            return v;
        };
        if (t <= 0x7f) return integer(t);
        if (t >= 0xe0) return integer(static_cast<int8_t>(t));
        if (t >= 0x80 && t <= 0x8f) return make_map(t & 0x0f, depth);
        if (t >= 0x90 && t <= 0x9f) return make_array(t & 0x0f, depth);
        if (t >= 0xa0 && t <= 0xbf) {
            v.kind = Value::string;
            v.s = str(t & 0x1f);
            return v;
        }
        switch (t) {
        case 0xc0: return v;
        case 0xc2:
        case 0xc3:
            v.kind = Value::boolean;
            v.b = t == 0xc3;
            return v;
        case 0xc4:
        case 0xd9:
            v.kind = Value::string;
            v.s = str(be(1));
            return v;
        case 0xc5:
        case 0xda:
            v.kind = Value::string;
            v.s = str(be(2));
            return v;
        case 0xc6:
        case 0xdb:
            v.kind = Value::string;
            v.s = str(be(4));
            return v;
        case 0xca: {
            uint32_t const bits = static_cast<uint32_t>(be(4));
            float x;
            std::memcpy(&x, &bits, 4);
            v.kind = Value::real;
            v.f = x;
            return v;
        }
        case 0xcb: {
            uint64_t const bits = be(8);
            std::memcpy(&v.f, &bits, 8);
            v.kind = Value::real;
            return v;
        }
        case 0xcc: return integer(static_cast<int64_t>(be(1)));
        case 0xcd: return integer(static_cast<int64_t>(be(2)));
        case 0xce: return integer(static_cast<int64_t>(be(4)));
        case 0xcf: {
            uint64_t const x = be(8);
            if (x > static_cast<uint64_t>(std::numeric_limits<int64_t>::max())) throw SerializationError("Integer value overflows the data type!\n");
            return integer(static_cast<int64_t>(x));
        }
        case 0xd0: return integer(static_cast<int8_t>(be(1)));
        case 0xd1: return integer(static_cast<int16_t>(be(2)));
        case 0xd2: return integer(static_cast<int32_t>(be(4)));
        case 0xd3: return integer(static_cast<int64_t>(be(8)));
        case 0xdc: return make_array(be(2), depth);
        case 0xdd: return make_array(be(4), depth);
        case 0xde: return make_map(be(2), depth);
        case 0xdf: return make_map(be(4), depth);
        default: fail("Error in parsing");
        }
    }
};

// ---- tree -> JSON text / msgpack bytes -------------------------------------------------------------------------------
void json_string(std::string& out, std::string const& s) {
    out.push_back('"');
    for (char const c : s) {
        switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\n': out += "\\n"; break;
        case '\r': out += "\\r"; break;
        case '\t': out += "\\t"; break;
        default:
            if (static_cast<unsigned char>(c) < 0x20) {
                char buf[8];
                std::snprintf(buf, sizeof(buf), "\\u%04x", c);
                out += buf;
            } else {
                out.push_back(c);
            }
        }
    }
    out.push_back('"');
}
void json_real(std::string& out, double v) { // shortest round-trip; a whole number keeps a ".0" so that it reads back as a real
    if (std::isinf(v)) {
        out += v > 0.0 ? "\"inf\"" : "\"-inf\"";
        return;
    }
    char buf[40];
    auto const r = std::to_chars(buf, buf + sizeof(buf), v);
    std::string_view const text(buf, static_cast<size_t>(r.ptr - buf));
    out += text;
    if (text.find_first_of(".en") == std::string_view::npos) out += ".0";
}
// indent < 0: compact.  Otherwise maps / arrays that hold maps or arrays are broken over lines down to the rows; a row (a map
// or list of plain values) stays on one line, as the reference's writer does below its max_indent_level.
void json_write(std::string& out, Value const& v, int indent, int level) {
    auto newline = [&](int lv) {
        if (indent < 0) return;
        out.push_back('\n');
        out.append(static_cast<size_t>(indent * lv), ' ');
    };
    auto holds_containers = [](Value const& x) {
        for (auto const& e : x.arr)
            if (e.kind == Value::array || e.kind == Value::map) return true;
        for (auto const& e : x.obj)
            if (e.second.kind == Value::array || e.second.kind == Value::map) return true;
        return false;
    };
    switch (v.kind) {
    case Value::nil: out += "null"; break;
    case Value::boolean: out += v.b ? "true" : "false"; break;
    case Value::integer: out += std::to_string(v.i); break;
    case Value::real: json_real(out, v.f); break;
    case Value::string: json_string(out, v.s); break;
    case Value::array: {
        bool const multi = indent >= 0 && holds_containers(v) && level < 4;
        out.push_back('[');
        for (size_t k = 0; k != v.arr.size(); ++k) {
            if (k != 0) out += (indent >= 0 && !multi) ? ", " : ",";
            if (multi) newline(level + 1);
            json_write(out, v.arr[k], multi ? indent : (indent >= 0 ? 1 << 20 : -1), level + 1);
        }
        if (multi && !v.arr.empty()) newline(level);
        out.push_back(']');
        break;
    }
    case Value::map: {
        bool const multi = indent >= 0 && indent < (1 << 20) && (level == 0 || holds_containers(v)) && level < 4;
        out.push_back('{');
        for (size_t k = 0; k != v.obj.size(); ++k) {
            if (k != 0) out += (indent >= 0 && !multi) ? ", " : ",";
            if (multi) newline(level + 1);
            json_string(out, v.obj[k].first);
            out += indent >= 0 ? ": " : ":";
            json_write(out, v.obj[k].second, multi ? indent : (indent >= 0 ? 1 << 20 : -1), level + 1);
        }
        if (multi && !v.obj.empty()) newline(level);
        out.push_back('}');
        break;
    }
    }
}

void mp_be(std::string& out, uint64_t v, int n) {
    for (int k = n - 1; k >= 0; --k) out.push_back(static_cast<char>((v >> (8 * k)) & 0xff));
}
void msgpack_write(std::string& out, Value const& v) {
    auto header = [&](size_t n, unsigned fix, unsigned fix_max, unsigned t16, unsigned t32) {
        if (n <= fix_max) {
            out.push_back(static_cast<char>(fix | n));
        } else if (n <= 0xffff) {
            out.push_back(static_cast<char>(t16));
            mp_be(out, n, 2);
        } else {
            out.push_back(static_cast<char>(t32));
            mp_be(out, n, 4);
        }
    };
    auto str = [&](std::string const& s) {
        if (s.size() <= 31) {
            out.push_back(static_cast<char>(0xa0 | s.size()));
        } else if (s.size() <= 0xff) {
            out.push_back(static_cast<char>(0xd9));
            mp_be(out, s.size(), 1);
        } else if (s.size() <= 0xffff) {
            out.push_back(static_cast<char>(0xda));
            mp_be(out, s.size(), 2);
        } else {
            out.push_back(static_cast<char>(0xdb));
            mp_be(out, s.size(), 4);
        }
        out += s;
    };
    switch (v.kind) {
    case Value::nil: out.push_back(static_cast<char>(0xc0)); break;
    case Value::boolean: out.push_back(static_cast<char>(v.b ? 0xc3 : 0xc2)); break;
    case Value::integer:
        if (v.i >= 0 && v.i <= 0x7f) {
            out.push_back(static_cast<char>(v.i));
        } else if (v.i < 0 && v.i >= -32) {
            out.push_back(static_cast<char>(v.i));
        } else if (v.i >= -128 && v.i <= 127) {
            out.push_back(static_cast<char>(0xd0));
            mp_be(out, static_cast<uint64_t>(v.i), 1);
        } else if (v.i >= -32768 && v.i <= 32767) {
            out.push_back(static_cast<char>(0xd1));
            mp_be(out, static_cast<uint64_t>(v.i), 2);
        } else if (v.i >= std::numeric_limits<int32_t>::min() && v.i <= std::numeric_limits<int32_t>::max()) {
            out.push_back(static_cast<char>(0xd2));
            mp_be(out, static_cast<uint64_t>(v.i), 4);
        } else {
            out.push_back(static_cast<char>(0xd3));
            mp_be(out, static_cast<uint64_t>(v.i), 8);
        }
        break;
    case Value::real: {
        uint64_t bits;
        std::memcpy(&bits, &v.f, 8);
        out.push_back(static_cast<char>(0xcb));
        mp_be(out, bits, 8);
        break;
    }
    case Value::string: str(v.s); break;
    case Value::array:
        header(v.arr.size(), 0x90, 15, 0xdc, 0xdd);
        for (auto const& e : v.arr) msgpack_write(out, e);
        break;
    case Value::map:
        header(v.obj.size(), 0x80, 15, 0xde, 0xdf);
        for (auto const& kv : v.obj) {
            str(kv.first);
            msgpack_write(out, kv.second);
        }
        break;
    }
}

// ---- tree <-> attribute values ------------------------------------------------------------------------------------------
constexpr int32_t kNaInt32 = std::numeric_limits<int32_t>::min();
constexpr int8_t kNaInt8 = std::numeric_limits<int8_t>::min();

double real_of(Value const& v, std::string const& where) {
    switch (v.kind) {
    case Value::real: return v.f;
    case Value::integer: return static_cast<double>(v.i);
    case Value::string:
        if (v.s == "inf" || v.s == "+inf") return std::numeric_limits<double>::infinity();
        if (v.s == "-inf") return -std::numeric_limits<double>::infinity();
        [[fallthrough]];
    default: throw SerializationError("Expect a number." + where);
    }
}
// writes one attribute of one element; nil leaves the null value that set_nan put there
void store_attribute(PGM_MetaAttribute const& a, char* dst, Value const& v, std::string const& where) {
    if (v.kind == Value::nil) return;
    switch (a.ctype) {
    case 0: {
        if (v.kind != Value::integer) throw SerializationError("Expect an integer." + where);
        if (v.i < std::numeric_limits<int32_t>::min() || v.i > std::numeric_limits<int32_t>::max()) throw SerializationError("Integer value overflows the data type!\n");
        int32_t const x = static_cast<int32_t>(v.i);
        std::memcpy(dst, &x, 4);
        break;
    }
    case 1: {
        if (v.kind != Value::integer && v.kind != Value::boolean) throw SerializationError("Expect an integer." + where);
        int64_t const raw = v.kind == Value::boolean ? (v.b ? 1 : 0) : v.i;
        if (raw < -128 || raw > 127) throw SerializationError("Integer value overflows the data type!\n");
        int8_t const x = static_cast<int8_t>(raw);
        std::memcpy(dst, &x, 1);
        break;
    }
    case 2: {
        double const x = real_of(v, where);
        std::memcpy(dst, &x, 8);
        break;
    }
    default: {
        if (v.kind != Value::array || v.arr.size() != 3) throw SerializationError("Expect an array of 3 numbers." + where);
        for (int k = 0; k != 3; ++k) {
            if (v.arr[static_cast<size_t>(k)].kind == Value::nil) continue;
            double const x = real_of(v.arr[static_cast<size_t>(k)], where);
            std::memcpy(dst + 8 * k, &x, 8);
        }
    }
    }
}
bool attribute_is_nan(PGM_MetaAttribute const& a, char const* src) {
    switch (a.ctype) {
    case 0: {
        int32_t x;
        std::memcpy(&x, src, 4);
        return x == kNaInt32;
    }
    case 1: {
        int8_t x;
        std::memcpy(&x, src, 1);
        return x == kNaInt8;
    }
    case 2: {
        double x;
        std::memcpy(&x, src, 8);
        return std::isnan(x);
    }
    default: {
        double x[3];
        std::memcpy(x, src, 24);
        return std::isnan(x[0]) && std::isnan(x[1]) && std::isnan(x[2]);
    }
    }
}
Value load_attribute(PGM_MetaAttribute const& a, char const* src) {
    Value v;
    if (attribute_is_nan(a, src)) return v;
    switch (a.ctype) {
    case 0: {
        int32_t x;
        std::memcpy(&x, src, 4);
        v.kind = Value::integer;
        v.i = x;
        break;
    }
    case 1: {
        int8_t x;
        std::memcpy(&x, src, 1);
        v.kind = Value::integer;
        v.i = x;
        break;
    }
    case 2:
        v.kind = Value::real;
        std::memcpy(&v.f, src, 8);
        break;
    default: {
        double x[3];
        std::memcpy(x, src, 24);
        v.kind = Value::array;
        for (double const e : x) {
            Value ev;
            if (!std::isnan(e)) {
                ev.kind = Value::real;
                ev.f = e;
            }
            v.arr.push_back(ev);
        }
    }
    }
    return v;
}

// one attribute of element `idx` of a dataset buffer, row-based or columnar (nullptr: the columnar buffer does not carry it)
char* element_attribute(DatasetBuffer const& b, PGM_MetaAttribute const& a, PGM_Idx idx) {
    if (!b.columnar()) return static_cast<char*>(b.data) + static_cast<size_t>(idx) * b.meta->size + a.offset;
    for (auto const& ab : b.attributes)
        if (ab.attribute == &a) return static_cast<char*>(ab.data) + static_cast<size_t>(idx) * a.size();
    return nullptr;
}

} // namespace

struct PGM_Deserializer {
    Value root;
    std::unique_ptr<PGM_WritableDataset> dataset;
    std::map<std::string, std::vector<PGM_MetaAttribute const*>> predefined; // component -> attributes of its compact rows
    std::vector<Value const*> scenarios;                                       // one map per scenario

    PGM_Deserializer(char const* data, size_t size, PGM_Idx format) {
        if (data == nullptr) throw std::invalid_argument("Received null pointer where a valid pointer was expected.\n");
        if (format == 0) {
            root = JsonReader{data, size}.parse_document();
        } else if (format == 1) {
            root = MsgpackReader{data, size}.parse_document();
        } else {
            throw SerializationError("Buffer data input not supported for serialization format " + std::to_string(format) + "\n");
        }
        if (root.kind != Value::map) throw SerializationError("Json root should be a map!\n");
        auto key = [&](char const* name) {
            Value const* v = root.find(name);
            if (v == nullptr) throw SerializationError(std::string("Key ") + name + " not found!\n");
            return v;
        };
        Value const* version = key("version");
        Value const* type = key("type");
        Value const* is_batch = key("is_batch");
        Value const* attributes = key("attributes");
        Value const* payload = key("data");
        if (version->kind != Value::string || type->kind != Value::string) throw SerializationError("Expect a string. Position of error: version / type\n");
        if (is_batch->kind != Value::boolean) throw SerializationError("Expect a boolean. Position of error: is_batch\n");
        if ((payload->kind == Value::map) == is_batch->b || (payload->kind != Value::map && payload->kind != Value::array)) {
            throw SerializationError("Map/Array type of data does not match is_batch!\n");
        }
        if (payload->kind == Value::map) {
            scenarios.push_back(payload);
        } else {
            for (auto const& s : payload->arr) {
                if (s.kind != Value::map) throw SerializationError("Expect a map. Position of error: data\n");
                scenarios.push_back(&s);
            }
        }
        PGM_Idx const batch_size = static_cast<PGM_Idx>(scenarios.size());
        dataset = std::make_unique<PGM_WritableDataset>(type->s.c_str(), is_batch->b ? 1 : 0, batch_size);
        PGM_MetaDataset const& meta = *dataset->meta;
        // components in the order of the meta data (the reference collects them in a set of meta-component pointers)
        std::set<PGM_MetaComponent const*> present;
        for (Value const* s : scenarios)
            for (auto const& kv : s->obj) {
                PGM_MetaComponent const* mc = meta.find(kv.first);
                if (mc == nullptr) throw SerializationError("Cannot find component with name: " + kv.first + "!\n Position of error: data/" + kv.first + "\n");
                if (kv.second.kind != Value::array) throw SerializationError("Expect an array. Position of error: data/" + kv.first + "\n");
                present.insert(mc);
            }
        for (PGM_MetaComponent const* mc : present) {
            std::vector<PGM_Idx> counter(static_cast<size_t>(batch_size), 0);
            bool only_lists = true;
            for (size_t s = 0; s != scenarios.size(); ++s) {
                Value const* rows = scenarios[s]->find(mc->name);
                if (rows == nullptr) continue;
                counter[s] = static_cast<PGM_Idx>(rows->arr.size());
                for (auto const& row : rows->arr) only_lists = only_lists && row.kind != Value::map;
            }
            bool uniform = true;
            for (size_t s = 1; s < counter.size(); ++s) uniform = uniform && counter[s] == counter[s - 1];
            PGM_Idx const per_scenario = !uniform ? -1 : (batch_size == 0 ? 0 : counter.front());
            PGM_Idx total = 0;
            for (PGM_Idx const c : counter) total += c;
            DatasetBuffer b{mc->name, per_scenario, per_scenario < 0 ? total : per_scenario * batch_size, nullptr, nullptr, mc, {}, false, {}};
            b.has_indications = only_lists;
            dataset->buffers.push_back(std::move(b));
        }
        if (attributes->kind != Value::map) throw SerializationError("Expect a map. Position of error: attributes\n");
        for (auto const& kv : attributes->obj) {
            PGM_MetaComponent const* mc = meta.find(kv.first);
            if (mc == nullptr) throw SerializationError("Cannot find component with name: " + kv.first + "!\n Position of error: attributes/" + kv.first + "\n");
            if (kv.second.kind != Value::array) throw SerializationError("Expect an array. Position of error: attributes/" + kv.first + "\n");
            std::vector<PGM_MetaAttribute const*> list;
            for (auto const& name : kv.second.arr) {
                if (name.kind != Value::string) throw SerializationError("Expect a string. Position of error: attributes/" + kv.first + "\n");
                PGM_MetaAttribute const* ma = mc->find(name.s);
                if (ma == nullptr) throw SerializationError("Cannot find attribute with name: " + name.s + "!\n Position of error: attributes/" + kv.first + "\n");
                list.push_back(ma);
            }
            if (DatasetBuffer* b = dataset->find(kv.first); b != nullptr && b->has_indications) b->indications = list;
            predefined[kv.first] = std::move(list);
        }
    }

    void parse_to_buffer() {
        for (auto& b : dataset->buffers) {
            bool const has_rows = b.data != nullptr;
            if (!has_rows && b.attributes.empty()) {
                if (b.total_elements == 0) continue;
                throw DatasetError("No buffer has been set for component '" + b.component + "'!\n");
            }
            if (b.elements_per_scenario < 0 && b.indptr == nullptr) throw DatasetError("For a non-uniform buffer, indptr should be supplied!\n");
            // null values everywhere first: absent attributes stay null
            if (has_rows) {
                b.meta->set_nan(b.data, 0, b.total_elements);
            } else {
                for (auto const& ab : b.attributes) {
                    PGM_MetaAttribute const& a = *ab.attribute;
                    char* dst = static_cast<char*>(ab.data);
                    for (PGM_Idx i = 0; i != b.total_elements; ++i) {
                        switch (a.ctype) {
                        case 0: std::memcpy(dst + 4 * i, &kNaInt32, 4); break;
                        case 1: std::memcpy(dst + i, &kNaInt8, 1); break;
                        case 2: {
                            double const nan = std::numeric_limits<double>::quiet_NaN();
                            std::memcpy(dst + 8 * i, &nan, 8);
                            break;
                        }
                        default: {
                            double const nan3[3] = {std::numeric_limits<double>::quiet_NaN(), std::numeric_limits<double>::quiet_NaN(),
                                                    std::numeric_limits<double>::quiet_NaN()};
                            std::memcpy(dst + 24 * i, nan3, 24);
                        }
                        }
                    }
                }
            }
            auto const pre = predefined.find(b.component);
            PGM_Idx offset = 0;
            auto* indptr = const_cast<PGM_Idx*>(b.indptr);
            if (indptr != nullptr) indptr[0] = 0;
            for (size_t s = 0; s != scenarios.size(); ++s) {
                Value const* rows = scenarios[s]->find(b.component);
                PGM_Idx const n = rows == nullptr ? 0 : static_cast<PGM_Idx>(rows->arr.size());
                for (PGM_Idx e = 0; e != n; ++e) {
                    Value const& row = rows->arr[static_cast<size_t>(e)];
                    std::string const where = " Position of error: data/" + std::to_string(s) + "/" + b.component + "/" + std::to_string(e) + "\n";
                    if (row.kind == Value::map) {
                        for (auto const& kv : row.obj) {
                            PGM_MetaAttribute const* ma = b.meta->find(kv.first);
                            if (ma == nullptr) continue; // unknown attributes are skipped (deserializer.hpp: parse_skip)
                            if (char* dst = element_attribute(b, *ma, offset + e); dst != nullptr) store_attribute(*ma, dst, kv.second, where);
                        }
                    } else if (row.kind == Value::array) {
                        if (pre == predefined.end() || pre->second.size() != row.arr.size()) {
                            throw SerializationError("An element of a list should have same length as the list of predefined attributes!\n" + where);
                        }
                        for (size_t k = 0; k != row.arr.size(); ++k) {
                            PGM_MetaAttribute const& ma = *pre->second[k];
                            if (char* dst = element_attribute(b, ma, offset + e); dst != nullptr) store_attribute(ma, dst, row.arr[k], where);
                        }
                    } else {
                        throw SerializationError("Expect a map or an array." + where);
                    }
                }
                offset += n;
                if (indptr != nullptr) indptr[s + 1] = offset;
            }
        }
    }
};

struct PGM_Serializer {
    Dataset const* dataset;
    PGM_Idx format;
    std::string buffer;

    PGM_Serializer(Dataset const& ds, PGM_Idx fmt) : dataset{&ds}, format{fmt} {
        if (fmt != 0 && fmt != 1) throw SerializationError("Unsupported serialization format: " + std::to_string(fmt) + "\n");
    }

    Value build(bool compact) const {
        Dataset const& ds = *dataset;
        Value root;
        root.kind = Value::map;
        auto text = [](std::string const& s) {
            Value v;
            v.kind = Value::string;
            v.s = s;
            return v;
        };
        Value flag;
        flag.kind = Value::boolean;
        flag.b = ds.is_batch;
        root.obj.emplace_back("version", text("1.0"));
        root.obj.emplace_back("type", text(ds.name));
        root.obj.emplace_back("is_batch", flag);
        // compact list: per component the attributes that carry a value somewhere in the buffer
        std::vector<std::vector<PGM_MetaAttribute const*>> kept(ds.buffers.size());
        Value attributes;
        attributes.kind = Value::map;
        for (size_t bi = 0; bi != ds.buffers.size(); ++bi) {
            DatasetBuffer const& b = ds.buffers[bi];
            for (int64_t k = 0; k != b.meta->n_attributes; ++k) {
                PGM_MetaAttribute const& a = b.meta->attributes[k];
                if (b.columnar() && element_attribute(b, a, 0) == nullptr && b.total_elements != 0) continue;
                bool any = false;
                for (PGM_Idx i = 0; i != b.total_elements && !any; ++i) any = !attribute_is_nan(a, element_attribute(b, a, i));
                if (any) kept[bi].push_back(&a);
            }
            if (compact && b.total_elements != 0) {
                Value names;
                names.kind = Value::array;
                for (auto const* a : kept[bi]) names.arr.push_back(text(a->name));
                attributes.obj.emplace_back(b.component, std::move(names));
            }
        }
        root.obj.emplace_back("attributes", std::move(attributes));
        auto scenario = [&](PGM_Idx s) {
            Value m;
            m.kind = Value::map;
            for (size_t bi = 0; bi != ds.buffers.size(); ++bi) {
                DatasetBuffer const& b = ds.buffers[bi];
                PGM_Idx const begin = b.elements_per_scenario < 0 ? b.indptr[s] : s * b.elements_per_scenario;
                PGM_Idx const end = b.elements_per_scenario < 0 ? b.indptr[s + 1] : (s + 1) * b.elements_per_scenario;
                if (begin == end) continue; // empty components are omitted
                Value rows;
                rows.kind = Value::array;
                for (PGM_Idx i = begin; i != end; ++i) {
                    Value row;
                    if (compact) {
                        row.kind = Value::array;
                        for (auto const* a : kept[bi]) row.arr.push_back(load_attribute(*a, element_attribute(b, *a, i)));
                    } else {
                        row.kind = Value::map;
                        for (auto const* a : kept[bi]) {
                            char const* src = element_attribute(b, *a, i);
                            if (!attribute_is_nan(*a, src)) row.obj.emplace_back(a->name, load_attribute(*a, src));
                        }
                    }
                    rows.arr.push_back(std::move(row));
                }
                m.obj.emplace_back(b.component, std::move(rows));
            }
            return m;
        };
        if (!ds.is_batch) {
            root.obj.emplace_back("data", scenario(0));
        } else {
            Value list;
            list.kind = Value::array;
            for (PGM_Idx s = 0; s != ds.batch_size; ++s) list.arr.push_back(scenario(s));
            root.obj.emplace_back("data", std::move(list));
        }
        return root;
    }
    std::string const& to_json(bool compact, PGM_Idx indent) {
        buffer.clear();
        json_write(buffer, build(compact), static_cast<int>(indent), 0);
        return buffer;
    }
    std::string const& to_binary(bool compact) {
        if (format == 0) return to_json(compact, -1);
        buffer.clear();
        msgpack_write(buffer, build(compact));
        return buffer;
    }
};

namespace {
// PGM_serialization_error instead of PGM_regular_error for everything the (de)serializer throws (handle.hpp:70-89)
template <class F> auto call_serialization(PGM_Handle* handle, F&& f) noexcept -> decltype(f()) {
    using R = decltype(f());
    try {
        clear(handle);
        return f();
    } catch (std::exception const& e) {
        if (handle != nullptr) {
            handle->err_code = PGM_serialization_error;
            handle->err_msg = e.what();
        }
    } catch (...) {
        if (handle != nullptr) {
            handle->err_code = PGM_serialization_error;
            handle->err_msg = "Unknown error!\n";
        }
    }
    if constexpr (!std::is_void_v<R>) return R{};
}
} // namespace

extern "C" {

PGM_Deserializer* PGM_create_deserializer_from_binary_buffer(PGM_Handle* handle, char const* data, PGM_Idx size, PGM_Idx serialization_format) {
    return call_serialization(handle, [&] { return new PGM_Deserializer{data, static_cast<size_t>(std::max<PGM_Idx>(size, 0)), serialization_format}; });
}
PGM_Deserializer* PGM_create_deserializer_from_null_terminated_string(PGM_Handle* handle, char const* data_string, PGM_Idx serialization_format) {
    return call_serialization(handle, [&] {
        if (data_string == nullptr) throw std::invalid_argument("Received null pointer where a valid pointer was expected.\n");
        if (serialization_format != 0) throw SerializationError("String data input not supported for serialization format " + std::to_string(serialization_format) + "\n");
        return new PGM_Deserializer{data_string, std::strlen(data_string), serialization_format};
    });
}
PGM_WritableDataset* PGM_deserializer_get_dataset(PGM_Handle* handle, PGM_Deserializer* deserializer) {
    return call_serialization(handle, [&] { return deref(deserializer).dataset.get(); });
}
void PGM_deserializer_parse_to_buffer(PGM_Handle* handle, PGM_Deserializer* deserializer) {
    call_serialization(handle, [&] { deref(deserializer).parse_to_buffer(); });
}
void PGM_destroy_deserializer(PGM_Deserializer* deserializer) { delete deserializer; }

PGM_Serializer* PGM_create_serializer(PGM_Handle* handle, PGM_ConstDataset const* dataset, PGM_Idx serialization_format) {
    return call_serialization(handle, [&] { return new PGM_Serializer{deref(dataset), serialization_format}; });
}
void PGM_serializer_get_to_binary_buffer(PGM_Handle* handle, PGM_Serializer* serializer, PGM_Idx use_compact_list, char const** data, PGM_Idx* size) {
    call_serialization(handle, [&] {
        std::string const& out = deref(serializer).to_binary(use_compact_list != 0);
        deref(data) = out.data();
        deref(size) = static_cast<PGM_Idx>(out.size());
    });
}
char const* PGM_serializer_get_to_zero_terminated_string(PGM_Handle* handle, PGM_Serializer* serializer, PGM_Idx use_compact_list, PGM_Idx indent) {
    return call_serialization(handle, [&]() -> char const* {
        PGM_Serializer& s = deref(serializer);
        if (s.format != 0) throw SerializationError("Serialization format " + std::to_string(s.format) + " does not support string output!\n");
        return s.to_json(use_compact_list != 0, indent).c_str();
    });
}
void PGM_destroy_serializer(PGM_Serializer* serializer) { delete serializer; }

// ---- writable dataset (dataset.h: the deserializer's dataset, buffers supplied by the caller) ----------------------
PGM_DatasetInfo const* PGM_dataset_writable_get_info(PGM_Handle* handle, PGM_WritableDataset const* dataset) {
    return call(handle, [&] { return reinterpret_cast<PGM_DatasetInfo const*>(static_cast<Dataset const*>(&deref(dataset))); });
}
void PGM_dataset_writable_set_buffer(PGM_Handle* handle, PGM_WritableDataset* dataset, char const* component, PGM_Idx* indptr, void* data) {
    call(handle, [&] {
        if (component == nullptr) throw std::invalid_argument("Received null pointer where a valid pointer was expected.\n");
        DatasetBuffer* b = deref(dataset).find(component);
        if (b == nullptr) throw DatasetError("Cannot find component '" + std::string(component) + "'!\n");
        if (b->elements_per_scenario < 0 && indptr == nullptr) throw DatasetError("For a non-uniform buffer, indptr should be supplied!\n");
        if (b->elements_per_scenario >= 0 && indptr != nullptr) throw DatasetError("For a uniform buffer, indptr should be nullptr!\n");
        b->indptr = indptr;
        b->data = data;
    });
}
void PGM_dataset_writable_set_attribute_buffer(PGM_Handle* handle, PGM_WritableDataset* dataset, char const* component, char const* attribute,
                                               void* data) {
    call(handle, [&] { deref(dataset).add_attribute_buffer(component, attribute, data); });
}
PGM_ConstDataset* PGM_create_dataset_const_from_writable(PGM_Handle* handle, PGM_WritableDataset const* writable_dataset) {
    return call(handle, [&] {
        Dataset const& src = deref(writable_dataset);
        auto* ds = new PGM_ConstDataset{src.name.c_str(), src.is_batch ? 1 : 0, src.batch_size};
        ds->buffers = src.buffers;
        return ds;
    });
}

} // extern "C"
