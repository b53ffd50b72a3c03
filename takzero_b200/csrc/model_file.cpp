// model_file.cpp -- reads the reference's model files without libtorch.
//
// The reference saves its network with tch `VarStore::save` (takzero/src/network/mod.rs:16-18,
// net6_simhash.rs:152-171; written by learn/src/main.rs:166,257, read by selfplay/src/main.rs:107,
// reanalyze/src/main.rs:93, tei/src/main.rs:50).  For a path that does not end in ".safetensors" that is
// libtorch's `torch::serialize::OutputArchive::write(name, tensor, /*buffer=*/true)` per variable followed by
// `save_to(path)`: a ZIP container (stored, 64-byte aligned entries) holding `<stem>/data.pkl` -- a protocol-2
// pickle of the module object whose state dict maps every variable name to
// `torch._utils._rebuild_tensor_v2(storage, offset, shape, stride, ...)` -- and one raw little-endian blob
// `<stem>/data/<key>` per storage.  `torch.save(state_dict)` files have the same layout (`data.pkl` is the dict),
// so both are accepted, as are safetensors files (what tch writes when the path ends in ".safetensors") and this
// repository's own TZW1 container (takzero_b200/weights.py::save_tzw).
//
// Host-only code: a ZIP central-directory walk, a small pickle stack machine (the opcodes libtorch and
// `torch.save` emit) and a strided gather to contiguous f32.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/takzero_b200.h"

namespace {

struct Fail : std::runtime_error {
    using std::runtime_error::runtime_error;
};

std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw Fail("cannot open " + path);
    std::vector<uint8_t> buf;
    if (std::fseek(f, 0, SEEK_END) == 0) {
        const long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        if (n > 0) buf.resize((size_t)n);
    }
    const size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    if (got != buf.size()) throw Fail(path + ": short read");
    return buf;
}

template <typename T>
T le(const uint8_t* p) {
    T v;
    std::memcpy(&v, p, sizeof(T));
    return v;
}

// ---- ZIP ---------------------------------------------------------------------------------------------

struct ZipEntry {
    std::string name;
    uint64_t offset = 0;  // of the data
    uint64_t size = 0;
};

std::vector<ZipEntry> zip_entries(const std::vector<uint8_t>& z) {
    const size_t n = z.size();
    if (n < 22) throw Fail("not a zip archive (too short)");
    size_t eocd = n;
    for (size_t back = 22; back <= std::min<size_t>(n, 22 + 65535); back++)
        if (le<uint32_t>(&z[n - back]) == 0x06054b50u) {
            eocd = n - back;
            break;
        }
    if (eocd == n) throw Fail("not a zip archive (no end-of-central-directory record)");
    uint64_t count = le<uint16_t>(&z[eocd + 10]);
    uint64_t cd_size = le<uint32_t>(&z[eocd + 12]);
    uint64_t cd_off = le<uint32_t>(&z[eocd + 16]);
    if (count == 0xffff || cd_size == 0xffffffffu || cd_off == 0xffffffffu) {  // zip64
        if (eocd < 20 || le<uint32_t>(&z[eocd - 20]) != 0x07064b50u) throw Fail("zip64 locator missing");
        const uint64_t e64 = le<uint64_t>(&z[eocd - 20 + 8]);
        if (e64 + 56 > n || le<uint32_t>(&z[e64]) != 0x06064b50u) throw Fail("zip64 end record missing");
        count = le<uint64_t>(&z[e64 + 32]);
        cd_size = le<uint64_t>(&z[e64 + 40]);
        cd_off = le<uint64_t>(&z[e64 + 48]);
    }
    if (cd_off + cd_size > n) throw Fail("zip central directory out of range");
    std::vector<ZipEntry> out;
    size_t p = (size_t)cd_off;
    for (uint64_t i = 0; i < count; i++) {
        if (p + 46 > n || le<uint32_t>(&z[p]) != 0x02014b50u) throw Fail("bad zip central directory entry");
        const uint16_t method = le<uint16_t>(&z[p + 10]);
        uint64_t csize = le<uint32_t>(&z[p + 20]), usize = le<uint32_t>(&z[p + 24]);
        const uint16_t nlen = le<uint16_t>(&z[p + 28]), xlen = le<uint16_t>(&z[p + 30]), clen = le<uint16_t>(&z[p + 32]);
        uint64_t lho = le<uint32_t>(&z[p + 42]);
        if (p + 46 + nlen + xlen + clen > n) throw Fail("bad zip central directory entry");
        ZipEntry e;
        e.name.assign((const char*)&z[p + 46], nlen);
        // zip64 extended information: the fields that are 0xffffffff above, in order
        for (size_t x = p + 46 + nlen, xe = x + xlen; x + 4 <= xe;) {
            const uint16_t id = le<uint16_t>(&z[x]), len = le<uint16_t>(&z[x + 2]);
            if (id == 1) {
                size_t q = x + 4;
                if (usize == 0xffffffffu && q + 8 <= xe) { usize = le<uint64_t>(&z[q]); q += 8; }
                if (csize == 0xffffffffu && q + 8 <= xe) { csize = le<uint64_t>(&z[q]); q += 8; }
                if (lho == 0xffffffffu && q + 8 <= xe) { lho = le<uint64_t>(&z[q]); q += 8; }
            }
            x += 4 + len;
        }
        p += 46 + nlen + xlen + clen;
        if (lho + 30 > n || le<uint32_t>(&z[lho]) != 0x04034b50u) throw Fail("bad zip local header for " + e.name);
        e.offset = lho + 30 + le<uint16_t>(&z[lho + 26]) + le<uint16_t>(&z[lho + 28]);
        e.size = usize;
        if (method != 0) {
            e.size = UINT64_MAX;  // compressed (libtorch deflates only code/*.py, which is never read here)
        } else if (e.offset + e.size > n) {
            throw Fail("zip entry out of range: " + e.name);
        }
        out.push_back(std::move(e));
    }
    return out;
}

bool ends_with(const std::string& s, const std::string& suffix) {
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

// ---- pickle ------------------------------------------------------------------------------------------

struct Val;
using VP = std::shared_ptr<Val>;
struct Val {
    enum Kind { NONE, BOOL, INT, FLOAT, STR, TUPLE, LIST, DICT, GLOBAL, STORAGE, TENSOR, OBJECT, MARK } kind = NONE;
    int64_t i = 0;           // BOOL / INT; TENSOR: storage offset; STORAGE: element count
    double f = 0;
    std::string s;           // STR; GLOBAL "module name"; STORAGE: key
    std::string dtype;       // STORAGE: torch storage class name
    std::vector<VP> items;   // TUPLE / LIST; DICT: key, value alternating; OBJECT: {class, state}; TENSOR: {storage}
    std::vector<int64_t> shape, stride;
};

VP mk(Val::Kind k) {
    auto v = std::make_shared<Val>();
    v->kind = k;
    return v;
}

struct Unpickler {
    const uint8_t* p;
    const uint8_t* end;
    std::vector<VP> stack;
    std::map<uint32_t, VP> memo;
    uint32_t next_memo = 0;

    const uint8_t* take(size_t k) {
        if ((size_t)(end - p) < k) throw Fail("pickle truncated");
        const uint8_t* q = p;
        p += k;
        return q;
    }
    VP pop() {
        if (stack.empty()) throw Fail("pickle stack underflow");
        VP v = stack.back();
        stack.pop_back();
        return v;
    }
    VP& top() {
        if (stack.empty()) throw Fail("pickle stack underflow");
        return stack.back();
    }
    std::vector<VP> pop_mark() {
        std::vector<VP> items;
        for (;;) {
            VP v = pop();
            if (v->kind == Val::MARK) break;
            items.push_back(v);
        }
        std::reverse(items.begin(), items.end());
        return items;
    }
    std::string line() {
        std::string s;
        for (;;) {
            const char c = (char)*take(1);
            if (c == '\n') return s;
            s += c;
        }
    }
    void push_str(size_t len) {
        VP v = mk(Val::STR);
        v->s.assign((const char*)take(len), len);
        stack.push_back(v);
    }
    void push_int(int64_t x) {
        VP v = mk(Val::INT);
        v->i = x;
        stack.push_back(v);
    }
    static int64_t as_int(const VP& v) {
        if (v->kind != Val::INT && v->kind != Val::BOOL) throw Fail("pickle: integer expected");
        return v->i;
    }
    static std::vector<int64_t> as_ints(const VP& v) {
        if (v->kind != Val::TUPLE && v->kind != Val::LIST) throw Fail("pickle: tuple of integers expected");
        std::vector<int64_t> out;
        for (const VP& x : v->items) out.push_back(as_int(x));
        return out;
    }

    VP reduce(const VP& fn, const VP& args) {
        if (fn->kind == Val::GLOBAL && args->kind == Val::TUPLE) {
            const std::string& g = fn->s;
            if (g == "torch._utils _rebuild_tensor_v2" || g == "torch._utils _rebuild_tensor") {
                if (args->items.size() < 4 || args->items[0]->kind != Val::STORAGE) throw Fail("pickle: bad _rebuild_tensor arguments");
                VP t = mk(Val::TENSOR);
                t->items.push_back(args->items[0]);
                t->i = as_int(args->items[1]);
                t->shape = as_ints(args->items[2]);
                t->stride = as_ints(args->items[3]);
                if (t->shape.size() != t->stride.size()) throw Fail("pickle: tensor shape / stride mismatch");
                return t;
            }
            if (g == "torch._utils _rebuild_parameter" || g == "torch._utils _rebuild_parameter_with_state") {
                if (args->items.empty()) throw Fail("pickle: bad _rebuild_parameter arguments");
                return args->items[0];
            }
            if (g == "collections OrderedDict") return mk(Val::DICT);
        }
        VP o = mk(Val::OBJECT);  // something this reader has no use for; kept opaque
        o->items = {fn, args};
        return o;
    }

    VP run() {
        for (;;) {
            const uint8_t op = *take(1);
            switch (op) {
                case 0x80: take(1); break;                       // PROTO
                case 0x95: take(8); break;                       // FRAME
                case '.': return pop();                          // STOP
                case '(': stack.push_back(mk(Val::MARK)); break;  // MARK
                case 'N': stack.push_back(mk(Val::NONE)); break;
                case 0x88: case 0x89: { VP v = mk(Val::BOOL); v->i = op == 0x88; stack.push_back(v); break; }
                case 'K': push_int(*take(1)); break;
                case 'M': push_int(le<uint16_t>(take(2))); break;
                case 'J': push_int(le<int32_t>(take(4))); break;
                case 0x8a: {  // LONG1
                    const size_t k = *take(1);
                    if (k > 8) throw Fail("pickle: integer too large");
                    const uint8_t* b = take(k);
                    uint64_t u = 0;
                    for (size_t j = 0; j < k; j++) u |= (uint64_t)b[j] << (8 * j);
                    if (k > 0 && k < 8 && (b[k - 1] & 0x80)) u |= ~uint64_t(0) << (8 * k);
                    push_int((int64_t)u);
                    break;
                }
                case 'G': {  // BINFLOAT, big endian
                    const uint8_t* b = take(8);
                    uint64_t u = 0;
                    for (int j = 0; j < 8; j++) u = (u << 8) | b[j];
                    VP v = mk(Val::FLOAT);
                    std::memcpy(&v->f, &u, 8);
                    stack.push_back(v);
                    break;
                }
                case 'X': case 'T': case 'B': push_str(le<uint32_t>(take(4))); break;  // BINUNICODE / BINSTRING / BINBYTES
                case 0x8c: case 'U': case 'C': push_str(*take(1)); break;              // SHORT_*
                case 0x8d: push_str((size_t)le<uint64_t>(take(8))); break;             // BINUNICODE8
                case 'c': {  // GLOBAL
                    VP v = mk(Val::GLOBAL);
                    const std::string module = line();
                    v->s = module + " " + line();
                    stack.push_back(v);
                    break;
                }
                case 0x93: {  // STACK_GLOBAL
                    VP name = pop(), module = pop();
                    VP v = mk(Val::GLOBAL);
                    v->s = module->s + " " + name->s;
                    stack.push_back(v);
                    break;
                }
                case 'q': memo[*take(1)] = top(); break;                 // BINPUT
                case 'r': memo[le<uint32_t>(take(4))] = top(); break;    // LONG_BINPUT
                case 0x94: memo[next_memo++] = top(); break;             // MEMOIZE
                case 'h': case 'j': {                                    // BINGET / LONG_BINGET
                    const uint32_t k = op == 'h' ? *take(1) : le<uint32_t>(take(4));
                    auto it = memo.find(k);
                    if (it == memo.end()) throw Fail("pickle: unknown memo key");
                    stack.push_back(it->second);
                    break;
                }
                case ')': stack.push_back(mk(Val::TUPLE)); break;
                case 't': { VP v = mk(Val::TUPLE); v->items = pop_mark(); stack.push_back(v); break; }
                case 0x85: case 0x86: case 0x87: {  // TUPLE1..3
                    VP v = mk(Val::TUPLE);
                    v->items.resize(op - 0x84);
                    for (size_t j = v->items.size(); j-- > 0;) v->items[j] = pop();
                    stack.push_back(v);
                    break;
                }
                case ']': stack.push_back(mk(Val::LIST)); break;
                case 'l': { VP v = mk(Val::LIST); v->items = pop_mark(); stack.push_back(v); break; }
                case 'a': { VP x = pop(); top()->items.push_back(x); break; }
                case 'e': { std::vector<VP> xs = pop_mark(); for (VP& x : xs) top()->items.push_back(x); break; }
                case '}': stack.push_back(mk(Val::DICT)); break;
                case 'd': { VP v = mk(Val::DICT); v->items = pop_mark(); stack.push_back(v); break; }
                case 's': { VP val = pop(), key = pop(); VP& d = top(); d->items.push_back(key); d->items.push_back(val); break; }
                case 'u': {
                    std::vector<VP> xs = pop_mark();
                    if (xs.size() % 2) throw Fail("pickle: odd SETITEMS");
                    VP& d = top();
                    if (d->kind != Val::DICT) throw Fail("pickle: SETITEMS on a non-dict");
                    for (VP& x : xs) d->items.push_back(x);
                    break;
                }
                case 'Q': {  // BINPERSID: ('storage', <storage class>, key, device, numel)
                    VP id = pop();
                    if (id->kind != Val::TUPLE || id->items.size() < 5 || id->items[0]->kind != Val::STR ||
                        id->items[0]->s != "storage" || id->items[1]->kind != Val::GLOBAL || id->items[2]->kind != Val::STR)
                        throw Fail("pickle: unsupported persistent id");
                    VP st = mk(Val::STORAGE);
                    st->dtype = id->items[1]->s;
                    st->s = id->items[2]->s;
                    st->i = as_int(id->items[4]);
                    stack.push_back(st);
                    break;
                }
                case 'R': { VP args = pop(), fn = pop(); stack.push_back(reduce(fn, args)); break; }
                case 0x81: {  // NEWOBJ
                    VP args = pop(), cls = pop();
                    VP o = mk(Val::OBJECT);
                    o->items = {cls, mk(Val::NONE)};
                    stack.push_back(o);
                    break;
                }
                case 'b': {  // BUILD
                    VP state = pop();
                    VP& o = top();
                    if (o->kind == Val::OBJECT) {
                        if (o->items.size() < 2) o->items.resize(2, mk(Val::NONE));
                        o->items[1] = state;
                    }  // tensors / dicts: backward-hook state, ignored
                    break;
                }
                case '0': pop(); break;
                case '2': stack.push_back(top()); break;
                default: {
                    char msg[64];
                    std::snprintf(msg, sizeof(msg), "pickle: unsupported opcode 0x%02x", op);
                    throw Fail(msg);
                }
            }
        }
    }
};

struct NamedTensor {
    std::string name;
    std::vector<int64_t> shape;
    std::vector<float> data;
};

float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

struct Archive {
    std::vector<uint8_t> bytes;
    std::vector<ZipEntry> entries;
    const ZipEntry* find_suffix(const std::string& a, const std::string& b) const {
        for (const ZipEntry& e : entries)
            if (e.name == b || ends_with(e.name, a)) return &e;
        return nullptr;
    }
};

void gather_tensor(const Archive& ar, const std::string& name, const Val& t, std::vector<NamedTensor>& out) {
    const Val& st = *t.items[0];
    int esize;
    const std::string& d = st.dtype;
    if (d == "torch FloatStorage") esize = 4;
    else if (d == "torch DoubleStorage") esize = 8;
    else if (d == "torch HalfStorage" || d == "torch BFloat16Storage") esize = 2;
    else if (d == "torch LongStorage") esize = 8;
    else if (d == "torch IntStorage") esize = 4;
    else return;  // not a numeric tensor this reader converts
    const ZipEntry* e = ar.find_suffix("/data/" + st.s, "data/" + st.s);
    if (!e || e->size == UINT64_MAX) throw Fail("storage " + st.s + " of " + name + " is missing from the archive");
    const uint64_t avail = e->size / (uint64_t)esize;
    size_t numel = 1;
    for (int64_t s : t.shape) {
        if (s < 0) throw Fail("negative dimension in " + name);
        numel *= (size_t)s;
        if (numel > (size_t(1) << 34)) throw Fail("tensor " + name + " is too large");
    }
    NamedTensor nt;
    nt.name = name;
    nt.shape = t.shape;
    nt.data.resize(numel);
    const uint8_t* base = ar.bytes.data() + e->offset;
    const size_t nd = t.shape.size();
    std::vector<int64_t> idx(nd, 0);
    int64_t pos = t.i;
    for (size_t k = 0; k < numel; k++) {
        if (pos < 0 || (uint64_t)pos >= avail) throw Fail("tensor " + name + " reaches outside its storage");
        const uint8_t* q = base + (size_t)pos * (size_t)esize;
        float v;
        if (d == "torch FloatStorage") v = le<float>(q);
        else if (d == "torch DoubleStorage") v = (float)le<double>(q);
        else if (d == "torch HalfStorage") v = half_to_float(le<uint16_t>(q));
        else if (d == "torch BFloat16Storage") { const uint32_t b = (uint32_t)le<uint16_t>(q) << 16; std::memcpy(&v, &b, 4); }
        else if (d == "torch LongStorage") v = (float)le<int64_t>(q);
        else v = (float)le<int32_t>(q);
        nt.data[k] = v;
        for (size_t a = nd; a-- > 0;) {  // odometer over the index, last dimension fastest
            pos += t.stride[a];
            if (++idx[a] < t.shape[a]) break;
            pos -= t.stride[a] * t.shape[a];
            idx[a] = 0;
        }
    }
    out.push_back(std::move(nt));
}

void collect(const Archive& ar, const std::string& prefix, const VP& v, std::vector<NamedTensor>& out, int depth) {
    if (depth > 8) return;
    if (v->kind == Val::OBJECT) {
        if (v->items.size() >= 2) collect(ar, prefix, v->items[1], out, depth + 1);
    } else if (v->kind == Val::DICT) {
        for (size_t k = 0; k + 1 < v->items.size(); k += 2) {
            const VP& key = v->items[k];
            const VP& val = v->items[k + 1];
            if (key->kind != Val::STR) continue;
            if (val->kind == Val::TENSOR) gather_tensor(ar, prefix + key->s, *val, out);
            else if (val->kind == Val::DICT || val->kind == Val::OBJECT) collect(ar, prefix + key->s + ".", val, out, depth + 1);
        }
    }
}

std::vector<NamedTensor> read_tzw(const std::vector<uint8_t>& b, const std::string& path) {
    size_t p = 8;
    const uint32_t count = le<uint32_t>(&b[4]);
    std::vector<NamedTensor> out;
    auto need = [&](size_t k) {
        if (p + k > b.size()) throw Fail(path + ": truncated");
    };
    for (uint32_t i = 0; i < count; i++) {
        need(4);
        const uint32_t len = le<uint32_t>(&b[p]);
        p += 4;
        need((size_t)len + 4);
        NamedTensor t;
        t.name.assign((const char*)&b[p], len);
        p += len;
        const uint32_t ndim = le<uint32_t>(&b[p]);
        p += 4;
        if (ndim > 8) throw Fail(path + ": bad rank");
        need(8 * (size_t)ndim);
        size_t numel = 1;
        for (uint32_t k = 0; k < ndim; k++) {
            const int64_t dim = le<int64_t>(&b[p + 8 * k]);
            if (dim < 0 || (numel *= (size_t)dim) > (size_t(1) << 34)) throw Fail(path + ": bad shape");
            t.shape.push_back(dim);
        }
        p += 8 * (size_t)ndim;
        need(4 * numel);
        t.data.resize(numel);
        std::memcpy(t.data.data(), &b[p], 4 * numel);
        p += 4 * numel;
        out.push_back(std::move(t));
    }
    return out;
}

// ---- safetensors (what tch's `VarStore::save` writes for a path ending in ".safetensors") -------------------------
// u64 little-endian header length, a JSON object {"name": {"dtype": "F32", "shape": [..], "data_offsets": [a, b]},
// ..., "__metadata__": {..}}, then the raw little-endian tensor bytes.  Only what that format needs of JSON.
struct Json {
    const char* p;
    const char* end;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    char peek() {
        ws();
        if (p >= end) throw Fail("safetensors header: unexpected end");
        return *p;
    }
    void expect(char c) {
        if (peek() != c) throw Fail(std::string("safetensors header: expected '") + c + "'");
        p++;
    }
    std::string string() {
        expect('"');
        std::string out;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                p++;
                if (*p == 'u') {  // names are ASCII in practice: keep \uXXXX below 0x80, replace the rest
                    if (p + 4 >= end) throw Fail("safetensors header: bad escape");
                    const unsigned v = (unsigned)std::strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                    out += v < 0x80 ? (char)v : '?';
                    p += 5;
                    continue;
                }
                out += *p == 'n' ? '\n' : *p == 't' ? '\t' : *p;
                p++;
                continue;
            }
            out += *p++;
        }
        expect('"');
        return out;
    }
    long long integer() {
        ws();
        char* e = nullptr;
        const long long v = std::strtoll(p, &e, 10);
        if (e == p) throw Fail("safetensors header: number expected");
        p = e;
        return v;
    }
    void skip_value() {  // any JSON value
        const char c = peek();
        if (c == '"') {
            string();
        } else if (c == '{' || c == '[') {
            const char close = c == '{' ? '}' : ']';
            p++;
            while (peek() != close) {
                if (c == '{') {
                    string();
                    expect(':');
                }
                skip_value();
                if (peek() == ',') p++;
            }
            p++;
        } else {
            while (p < end && *p != ',' && *p != '}' && *p != ']') p++;
        }
    }
};

std::vector<NamedTensor> read_safetensors(const std::vector<uint8_t>& b, const std::string& path) {
    if (b.size() < 8) throw Fail(path + ": truncated");
    const uint64_t hlen = le<uint64_t>(b.data());
    if (hlen > b.size() - 8) throw Fail(path + ": safetensors header out of range");
    const uint8_t* data = b.data() + 8 + hlen;
    const uint64_t data_len = b.size() - 8 - hlen;
    Json j{(const char*)b.data() + 8, (const char*)b.data() + 8 + hlen};
    struct Entry {
        NamedTensor t;
        uint64_t begin;
    };
    std::vector<Entry> entries;
    j.expect('{');
    while (j.peek() != '}') {
        const std::string name = j.string();
        j.expect(':');
        if (name == "__metadata__") {
            j.skip_value();
        } else {
            std::string dtype;
            std::vector<int64_t> shape;
            long long begin = -1, stop = -1;
            j.expect('{');
            while (j.peek() != '}') {
                const std::string key = j.string();
                j.expect(':');
                if (key == "dtype") {
                    dtype = j.string();
                } else if (key == "shape") {
                    j.expect('[');
                    while (j.peek() != ']') {
                        shape.push_back(j.integer());
                        if (j.peek() == ',') j.p++;
                    }
                    j.expect(']');
                } else if (key == "data_offsets") {
                    j.expect('[');
                    begin = j.integer();
                    j.expect(',');
                    stop = j.integer();
                    j.expect(']');
                } else {
                    j.skip_value();
                }
                if (j.peek() == ',') j.p++;
            }
            j.expect('}');
            const int esize = dtype == "F32" || dtype == "I32" ? 4 : dtype == "F64" || dtype == "I64" ? 8
                              : dtype == "F16" || dtype == "BF16" ? 2 : 0;
            size_t numel = 1;
            for (int64_t d : shape) {
                if (d < 0 || (numel *= (size_t)d) > (size_t(1) << 34)) throw Fail(path + ": bad shape of " + name);
            }
            if (esize != 0) {
                if (begin < 0 || stop < begin || (uint64_t)stop > data_len || (uint64_t)(stop - begin) != numel * (uint64_t)esize)
                    throw Fail(path + ": data_offsets of " + name + " do not match its shape");
                Entry e;
                e.t.name = name;
                e.t.shape = shape;
                e.t.data.resize(numel);
                const uint8_t* q = data + begin;
                for (size_t k = 0; k < numel; k++, q += esize) {
                    float v;
                    if (dtype == "F32") v = le<float>(q);
                    else if (dtype == "F64") v = (float)le<double>(q);
                    else if (dtype == "F16") v = half_to_float(le<uint16_t>(q));
                    else if (dtype == "BF16") { const uint32_t w = (uint32_t)le<uint16_t>(q) << 16; std::memcpy(&v, &w, 4); }
                    else if (dtype == "I64") v = (float)le<int64_t>(q);
                    else v = (float)le<int32_t>(q);
                    e.t.data[k] = v;
                }
                e.begin = (uint64_t)begin;
                entries.push_back(std::move(e));
            }
        }
        if (j.peek() == ',') j.p++;
    }
    // file order = order of the data, as a reader of the raw bytes sees the tensors
    std::stable_sort(entries.begin(), entries.end(), [](const Entry& a, const Entry& c) { return a.begin < c.begin; });
    std::vector<NamedTensor> out;
    for (Entry& e : entries) out.push_back(std::move(e.t));
    if (out.empty()) throw Fail(path + ": no tensors found");
    return out;
}

std::vector<NamedTensor> read_model(const std::string& path) {
    Archive ar;
    ar.bytes = read_file(path);
    if (ar.bytes.size() >= 8 && std::memcmp(ar.bytes.data(), "TZW1", 4) == 0) return read_tzw(ar.bytes, path);
    if (ar.bytes.size() >= 10 && ar.bytes[8] == '{' && le<uint64_t>(ar.bytes.data()) <= ar.bytes.size() - 8)
        return read_safetensors(ar.bytes, path);
    if (ar.bytes.size() < 4 || le<uint32_t>(ar.bytes.data()) != 0x04034b50u)
        throw Fail(path + ": neither a libtorch archive (.ot / .pt zip), a safetensors file nor a TZW1 file");
    ar.entries = zip_entries(ar.bytes);
    const ZipEntry* pkl = ar.find_suffix("/data.pkl", "data.pkl");
    if (!pkl || pkl->size == UINT64_MAX) throw Fail(path + ": no data.pkl in the archive");
    Unpickler u;
    u.p = ar.bytes.data() + pkl->offset;
    u.end = u.p + pkl->size;
    const VP root = u.run();
    std::vector<NamedTensor> out;
    collect(ar, "", root, out, 0);
    if (out.empty()) throw Fail(path + ": no tensors found");
    return out;
}

// tch `VarStore` names -> the names tz_set_weights documents.  Both `SmallBlock`s of a `ResidualBlock` are
// created on the SAME path (network/residual.rs:52-54), so tch registers the first one's variables as
// `core.res_block_B.{conv2d,batch_norm}.X` and, the name being taken, the second one's as
// `core.res_block_B.{conv2d,batch_norm}.X__K` (K = number of variables in the store at that moment; tch
// `nn::Path::add`).  Files written from Python modules use '|'-free dotted names already.
std::string canonical_name(std::string name) {
    std::replace(name.begin(), name.end(), '|', '.');
    const std::string pre = "core.res_block_";
    if (name.compare(0, pre.size(), pre) != 0) return name;
    size_t p = pre.size();
    while (p < name.size() && name[p] >= '0' && name[p] <= '9') p++;
    if (p == pre.size() || p >= name.size() || name[p] != '.') return name;
    const std::string rest = name.substr(p + 1);
    if (rest.compare(0, 7, "conv2d.") != 0 && rest.compare(0, 11, "batch_norm.") != 0) return name;  // already ".0." / ".1."
    const size_t us = rest.rfind("__");
    if (us == std::string::npos) return name.substr(0, p) + ".0." + rest;
    for (size_t k = us + 2; k < rest.size(); k++)
        if (rest[k] < '0' || rest[k] > '9') return name;
    return name.substr(0, p) + ".1." + rest.substr(0, us);
}

}  // namespace

void tz_internal_set_error(const char* msg);  // api.cu: the message tz_last_error() returns

static int model_fail(int code, const std::string& msg) {
    tz_internal_set_error(msg.c_str());
    return code;
}

extern "C" TZ_API int tz_read_model_file(const char* path, tz_model_tensor_fn fn, void* ctx) {
    if (!path || !fn) return model_fail(TZ_EINVAL, "null argument");
    try {
        const std::vector<NamedTensor> ts = read_model(path);
        for (const NamedTensor& t : ts) {
            const std::string name = canonical_name(t.name);
            fn(ctx, name.c_str(), t.name.c_str(), t.data.data(), t.shape.data(), (int)t.shape.size());
        }
        return (int)ts.size();
    } catch (const std::exception& e) {
        return model_fail(TZ_EINVAL, e.what());
    }
}

// Net::load (network/mod.rs:20-27, net6_simhash.rs:164-181): VarStore::load of `path`, then the SimHash set from
// the sidecar `bitvec.bin` in the same directory.  A missing sidecar is an error, as in the reference (with an empty
// set every local uncertainty would silently become 4.0); allow_missing_set = 1 takes it as the empty set of a freshly
// initialised network instead.  Networks without `simhash_matrix` / `lcghash_init` (net5) skip the novelty part.
extern "C" TZ_API int tz_load_model_ex(tz_handle* h, const char* path, int allow_missing_set) {
    if (!h || !path) return model_fail(TZ_EINVAL, "null argument");
    std::vector<NamedTensor> ts;
    try {
        ts = read_model(path);
    } catch (const std::exception& e) {
        return model_fail(TZ_EINVAL, e.what());
    }
    std::vector<std::string> names;
    for (const NamedTensor& t : ts) names.push_back(canonical_name(t.name));
    std::vector<tz_tensor_t> args;
    const NamedTensor* simhash = nullptr;
    const NamedTensor* lcghash = nullptr;
    for (size_t i = 0; i < ts.size(); i++) {
        if (names[i] == "simhash_matrix") {
            simhash = &ts[i];
            continue;
        }
        if (names[i] == "lcghash_init") {  // net4_lcghash.rs:131-137
            lcghash = &ts[i];
            continue;
        }
        if (ts[i].shape.size() > 4) continue;
        args.push_back(tz_tensor_t{names[i].c_str(), ts[i].data.data(), ts[i].shape.data(), (int)ts[i].shape.size()});
    }
    // the sidecar is read (or found missing) before the current model is touched: a failed load keeps the old one
    std::vector<uint8_t> bits;
    if (simhash || lcghash) {
        std::string dir = path;
        const size_t slash = dir.find_last_of('/');
        dir = slash == std::string::npos ? std::string() : dir.substr(0, slash + 1);
        FILE* f = std::fopen((dir + "bitvec.bin").c_str(), "rb");
        if (f) {
            bits.resize(size_t(1) << 29);
            const size_t got = std::fread(bits.data(), 1, bits.size(), f);
            std::fclose(f);
            if (got != bits.size()) return model_fail(TZ_EINVAL, dir + "bitvec.bin: expected 2^29 bytes");
        } else if (!allow_missing_set) {
            return model_fail(TZ_EINVAL, dir + "bitvec.bin is missing (the novelty set saved next to the model, "
                              "net6_simhash.rs:164-181); tz_load_model_ex(.., 1) loads the model with an empty set");
        }
    }
    int rc = tz_set_weights(h, args.data(), (int)args.size());
    if (rc != TZ_OK) return rc;
    if (simhash || lcghash) {
        rc = lcghash ? tz_set_lcghash(h, lcghash->data.data(), bits.empty() ? nullptr : bits.data())
                     : tz_set_simhash(h, simhash->data.data(), bits.empty() ? nullptr : bits.data());
        if (rc != TZ_OK) return rc;
    }
    return TZ_OK;
}

extern "C" TZ_API int tz_load_model(tz_handle* h, const char* path) { return tz_load_model_ex(h, path, 0); }
