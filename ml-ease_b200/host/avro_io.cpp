// avro_io.cpp -- see avro_io.hpp
#include "avro_io.hpp"

#include <dirent.h>
#include <sched.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <future>
#include <thread>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>

namespace mlease_host {

// ------------------------------------------------------------------------------------------ JSON
namespace {
struct JP {
  const std::string& t;
  size_t i = 0;
  explicit JP(const std::string& s) : t(s) {}
  void ws() { while (i < t.size() && (t[i] == ' ' || t[i] == '\n' || t[i] == '\t' || t[i] == '\r')) i++; }
  [[noreturn]] void err(const char* m) { throw std::runtime_error(std::string("json: ") + m + " at " + std::to_string(i)); }
  Json val() {
    ws();
    if (i >= t.size()) err("eof");
    Json j;
    char c = t[i];
    if (c == '{') {
      j.kind = Json::Obj; i++; ws();
      if (t[i] == '}') { i++; return j; }
      while (true) {
        ws(); Json k = val(); if (k.kind != Json::Str) err("key");
        ws(); if (t[i] != ':') err(":"); i++;
        j.obj.emplace_back(k.str, val());
        ws(); if (t[i] == ',') { i++; continue; }
        if (t[i] == '}') { i++; break; }
        err("obj");
      }
    } else if (c == '[') {
      j.kind = Json::Arr; i++; ws();
      if (t[i] == ']') { i++; return j; }
      while (true) {
        j.arr.push_back(val());
        ws(); if (t[i] == ',') { i++; continue; }
        if (t[i] == ']') { i++; break; }
        err("arr");
      }
    } else if (c == '"') {
      j.kind = Json::Str; i++;
      while (i < t.size() && t[i] != '"') {
        if (t[i] == '\\') {
          i++;
          char e = t[i++];
          switch (e) {
            case 'n': j.str.push_back('\n'); break;
            case 't': j.str.push_back('\t'); break;
            case 'r': j.str.push_back('\r'); break;
            case 'b': j.str.push_back('\b'); break;
            case 'f': j.str.push_back('\f'); break;
            case 'u': {
              unsigned cp = std::stoul(t.substr(i, 4), nullptr, 16); i += 4;
              if (cp < 0x80) j.str.push_back((char)cp);
              else if (cp < 0x800) { j.str.push_back((char)(0xC0 | (cp >> 6))); j.str.push_back((char)(0x80 | (cp & 0x3F))); }
              else { j.str.push_back((char)(0xE0 | (cp >> 12))); j.str.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); j.str.push_back((char)(0x80 | (cp & 0x3F))); }
              break;
            }
            default: j.str.push_back(e);
          }
        } else j.str.push_back(t[i++]);
      }
      i++;
    } else if (!t.compare(i, 4, "true")) { j.kind = Json::Bool; j.b = true; i += 4; }
    else if (!t.compare(i, 5, "false")) { j.kind = Json::Bool; j.b = false; i += 5; }
    else if (!t.compare(i, 4, "null")) { j.kind = Json::Null; i += 4; }
    else {
      size_t st = i;
      while (i < t.size() && (isdigit((unsigned char)t[i]) || t[i] == '-' || t[i] == '+' || t[i] == '.' || t[i] == 'e' || t[i] == 'E')) i++;
      if (st == i) err("value");
      j.kind = Json::Num; j.num = std::stod(t.substr(st, i - st));
    }
    return j;
  }
};
void dump(const Json& j, std::string& o) {
  switch (j.kind) {
    case Json::Null: o += "null"; break;
    case Json::Bool: o += j.b ? "true" : "false"; break;
    case Json::Num: { char b[64]; if (j.num == (long long)j.num) snprintf(b, 64, "%lld", (long long)j.num); else snprintf(b, 64, "%.17g", j.num); o += b; break; }
    case Json::Str:
      o.push_back('"');
      for (char c : j.str) { if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back(c); } else if (c == '\n') o += "\\n"; else o.push_back(c); }
      o.push_back('"');
      break;
    case Json::Arr:
      o.push_back('[');
      for (size_t k = 0; k < j.arr.size(); k++) { if (k) o.push_back(','); dump(j.arr[k], o); }
      o.push_back(']');
      break;
    case Json::Obj:
      o.push_back('{');
      for (size_t k = 0; k < j.obj.size(); k++) { if (k) o.push_back(','); Json s; s.kind = Json::Str; s.str = j.obj[k].first; dump(s, o); o.push_back(':'); dump(j.obj[k].second, o); }
      o.push_back('}');
      break;
  }
}
Json jstr(const std::string& s) { Json j; j.kind = Json::Str; j.str = s; return j; }
}  // namespace

Json json_parse(const std::string& text) { JP p(text); return p.val(); }
std::string json_dump(const Json& j) { std::string o; dump(j, o); return o; }

// ------------------------------------------------------------------------------------------ schema
SchemaP schema_from_json(const Json& j, std::map<std::string, SchemaP>& named) {
  auto prim = [&](const std::string& n) -> SchemaP {
    static const std::pair<const char*, Schema::Type> P[] = {{"null", Schema::Null}, {"boolean", Schema::Boolean}, {"int", Schema::Int},
        {"long", Schema::Long}, {"float", Schema::Float}, {"double", Schema::Double}, {"string", Schema::String}, {"bytes", Schema::Bytes}};
    for (auto& p : P) if (n == p.first) { auto s = std::make_shared<Schema>(); s->type = p.second; return s; }
    auto it = named.find(n);
    if (it != named.end()) return it->second;
    size_t dot = n.rfind('.');
    if (dot != std::string::npos) { it = named.find(n.substr(dot + 1)); if (it != named.end()) return it->second; }
    throw std::runtime_error("avro: unknown type " + n);
  };
  if (j.kind == Json::Str) return prim(j.str);
  if (j.kind == Json::Arr) {
    auto s = std::make_shared<Schema>(); s->type = Schema::Union;
    for (auto& b : j.arr) s->branches.push_back(schema_from_json(b, named));
    return s;
  }
  if (j.kind != Json::Obj) throw std::runtime_error("avro: bad schema");
  const Json* t = j.get("type");
  if (!t) throw std::runtime_error("avro: schema without type");
  if (t->kind != Json::Str) return schema_from_json(*t, named);
  const std::string& ty = t->str;
  if (ty == "record") {
    auto s = std::make_shared<Schema>(); s->type = Schema::Record;
    if (const Json* n = j.get("name")) s->name = n->str;
    named[s->name] = s;
    if (const Json* f = j.get("fields"))
      for (auto& fj : f->arr) s->fields.emplace_back(fj.get("name")->str, schema_from_json(*fj.get("type"), named));
    return s;
  }
  if (ty == "array") { auto s = std::make_shared<Schema>(); s->type = Schema::Array; s->items = schema_from_json(*j.get("items"), named); return s; }
  if (ty == "map") { auto s = std::make_shared<Schema>(); s->type = Schema::Map; s->items = schema_from_json(*j.get("values"), named); return s; }
  if (ty == "enum") {
    auto s = std::make_shared<Schema>(); s->type = Schema::Enum; s->name = j.get("name")->str;
    for (auto& e : j.get("symbols")->arr) s->symbols.push_back(e.str);
    named[s->name] = s; return s;
  }
  if (ty == "fixed") { auto s = std::make_shared<Schema>(); s->type = Schema::Fixed; s->name = j.get("name")->str; s->fixed_size = (int)j.get("size")->num; named[s->name] = s; return s; }
  return prim(ty);
}
SchemaP schema_parse(const std::string& text) { std::map<std::string, SchemaP> named; return schema_from_json(json_parse(text), named); }

Json schema_to_json(const SchemaP& s, std::map<std::string, bool>& emitted) {
  static const char* N[] = {"null", "boolean", "int", "long", "float", "double", "string", "bytes"};
  Json j;
  switch (s->type) {
    case Schema::Record: {
      if (emitted[s->name]) return jstr(s->name);
      emitted[s->name] = true;
      j.kind = Json::Obj;
      j.obj.emplace_back("type", jstr("record")); j.obj.emplace_back("name", jstr(s->name));
      Json f; f.kind = Json::Arr;
      for (auto& fd : s->fields) { Json o; o.kind = Json::Obj; o.obj.emplace_back("name", jstr(fd.first)); o.obj.emplace_back("type", schema_to_json(fd.second, emitted)); f.arr.push_back(o); }
      j.obj.emplace_back("fields", f);
      return j;
    }
    case Schema::Array: j.kind = Json::Obj; j.obj.emplace_back("type", jstr("array")); j.obj.emplace_back("items", schema_to_json(s->items, emitted)); return j;
    case Schema::Map: j.kind = Json::Obj; j.obj.emplace_back("type", jstr("map")); j.obj.emplace_back("values", schema_to_json(s->items, emitted)); return j;
    case Schema::Union: j.kind = Json::Arr; for (auto& b : s->branches) j.arr.push_back(schema_to_json(b, emitted)); return j;
    case Schema::Enum: { j.kind = Json::Obj; j.obj.emplace_back("type", jstr("enum")); j.obj.emplace_back("name", jstr(s->name)); Json a; a.kind = Json::Arr; for (auto& e : s->symbols) a.arr.push_back(jstr(e)); j.obj.emplace_back("symbols", a); return j; }
    case Schema::Fixed: { j.kind = Json::Obj; j.obj.emplace_back("type", jstr("fixed")); j.obj.emplace_back("name", jstr(s->name)); Json n; n.kind = Json::Num; n.num = s->fixed_size; j.obj.emplace_back("size", n); return j; }
    default: return jstr(N[s->type]);
  }
}

// utils/Util.java:377-417: a union collapses to its first non-null branch, recursively through arrays and records
SchemaP schema_remove_union(const SchemaP& s) {
  if (s->type == Schema::Union) {
    for (auto& b : s->branches) if (b->type != Schema::Null) return schema_remove_union(b);
    return s;
  }
  if (s->type == Schema::Array) { auto o = std::make_shared<Schema>(*s); o->items = schema_remove_union(s->items); return o; }
  if (s->type == Schema::Record) { auto o = std::make_shared<Schema>(*s); for (auto& f : o->fields) f.second = schema_remove_union(f.second); return o; }
  return s;
}

// ------------------------------------------------------------------------------------------ binary codec
namespace {
int64_t rd_long(const std::string& b, size_t& p) {
  uint64_t acc = 0; int sh = 0;
  while (true) {
    if (p >= b.size()) throw std::runtime_error("avro: truncated varint");
    uint8_t c = (uint8_t)b[p++];
    if (sh > 63) throw std::runtime_error("avro: varint longer than 10 bytes");
    acc |= (uint64_t)(c & 0x7F) << sh;
    if (!(c & 0x80)) break;
    sh += 7;
  }
  return (int64_t)(acc >> 1) ^ -(int64_t)(acc & 1);
}
void wr_long(std::string& o, int64_t v) {
  uint64_t z = ((uint64_t)v << 1) ^ (uint64_t)(v >> 63);
  while (z & ~0x7FULL) { o.push_back((char)((z & 0x7F) | 0x80)); z >>= 7; }
  o.push_back((char)z);
}
// every read of file content is bounds-checked: a truncated or corrupt data file is an error, never an out-of-bounds read
inline void need(const std::string& b, size_t p, int64_t n) {
  if (n < 0 || (uint64_t)n > b.size() || p > b.size() - (size_t)n) throw std::runtime_error("avro: truncated or corrupt block (value runs past the end of the data)");
}
void decode(const std::string& b, size_t& p, const Schema& s, Value& v) {
  v.type = s.type; v.items.clear(); v.s.clear();
  switch (s.type) {
    case Schema::Null: break;
    case Schema::Boolean: need(b, p, 1); v.i = b[p++] != 0; break;
    case Schema::Int: case Schema::Long: v.i = rd_long(b, p); break;
    case Schema::Float: { need(b, p, 4); float f; memcpy(&f, &b[p], 4); p += 4; v.d = f; break; }
    case Schema::Double: { need(b, p, 8); double d; memcpy(&d, &b[p], 8); p += 8; v.d = d; break; }
    case Schema::String: case Schema::Bytes: { int64_t n = rd_long(b, p); need(b, p, n); v.s.assign(b, p, (size_t)n); p += (size_t)n; break; }
    case Schema::Fixed: need(b, p, s.fixed_size); v.s.assign(b, p, (size_t)s.fixed_size); p += s.fixed_size; break;
    case Schema::Enum: v.i = rd_long(b, p); break;
    case Schema::Union: { int64_t br = rd_long(b, p); if (br < 0 || br >= (int64_t)s.branches.size()) throw std::runtime_error("avro: bad union branch"); decode(b, p, *s.branches[br], v); break; }
    case Schema::Record: v.items.resize(s.fields.size()); for (size_t k = 0; k < s.fields.size(); k++) decode(b, p, *s.fields[k].second, v.items[k]); v.type = Schema::Record; break;
    case Schema::Array:
      while (true) {
        int64_t n = rd_long(b, p);
        if (n == 0) break;
        if (n < 0) { n = -n; rd_long(b, p); }
        for (int64_t k = 0; k < n; k++) { v.items.emplace_back(); decode(b, p, *s.items, v.items.back()); }
      }
      v.type = Schema::Array;
      break;
    case Schema::Map:
      while (true) {
        int64_t n = rd_long(b, p);
        if (n == 0) break;
        if (n < 0) { n = -n; rd_long(b, p); }
        for (int64_t k = 0; k < n; k++) {
          int64_t l = rd_long(b, p);
          need(b, p, l);
          std::string key(b, p, (size_t)l); p += (size_t)l;
          v.items.emplace_back(); decode(b, p, *s.items, v.items.back());
          v.items.back().map_key = key;   // kept apart from the value (a string value has its own .s)
        }
      }
      v.type = Schema::Map;
      break;
  }
}
bool branch_matches(const Schema& br, const Value& v) {
  if (v.is_null()) return br.type == Schema::Null;
  switch (br.type) {
    case Schema::Int: case Schema::Long: return v.type == Schema::Int || v.type == Schema::Long;
    case Schema::Float: case Schema::Double: return v.type == Schema::Float || v.type == Schema::Double || v.type == Schema::Int || v.type == Schema::Long;
    case Schema::String: case Schema::Bytes: return v.type == Schema::String || v.type == Schema::Bytes;
    default: return br.type == v.type;
  }
}
void encode(std::string& o, const Schema& s, const Value& v) {
  switch (s.type) {
    case Schema::Null: break;
    case Schema::Boolean: o.push_back(v.i ? 1 : 0); break;
    case Schema::Int: case Schema::Long: case Schema::Enum:
      if (v.is_null()) throw std::runtime_error("avro: null for non-nullable int field");
      wr_long(o, (v.type == Schema::Float || v.type == Schema::Double) ? (int64_t)v.d : v.i); break;
    case Schema::Float: { if (v.is_null()) throw std::runtime_error("avro: null for non-nullable float field"); float f = (v.type == Schema::Int || v.type == Schema::Long) ? (float)v.i : (float)v.d; o.append((const char*)&f, 4); break; }
    case Schema::Double: { if (v.is_null()) throw std::runtime_error("avro: null for non-nullable double field"); double d = (v.type == Schema::Int || v.type == Schema::Long) ? (double)v.i : v.d; o.append((const char*)&d, 8); break; }
    case Schema::String: case Schema::Bytes: if (v.is_null()) throw std::runtime_error("avro: null for non-nullable string field"); wr_long(o, (int64_t)v.s.size()); o += v.s; break;
    case Schema::Fixed: o += v.s; break;
    case Schema::Union: {
      for (size_t k = 0; k < s.branches.size(); k++)
        if (branch_matches(*s.branches[k], v)) { wr_long(o, (int64_t)k); encode(o, *s.branches[k], v); return; }
      throw std::runtime_error("avro: value matches no union branch");
    }
    case Schema::Record:
      if (v.items.size() != s.fields.size()) throw std::runtime_error("avro: record arity mismatch for " + s.name);
      for (size_t k = 0; k < s.fields.size(); k++) encode(o, *s.fields[k].second, v.items[k]);
      break;
    case Schema::Array:
      if (!v.items.empty()) { wr_long(o, (int64_t)v.items.size()); for (auto& e : v.items) encode(o, *s.items, e); }
      wr_long(o, 0);
      break;
    case Schema::Map:
      if (!v.items.empty()) { wr_long(o, (int64_t)v.items.size()); for (auto& e : v.items) { wr_long(o, (int64_t)e.map_key.size()); o += e.map_key; encode(o, *s.items, e); } }
      wr_long(o, 0);
      break;
  }
}
std::string inflate_raw(const char* in, size_t in_size) {
  z_stream zs; memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("zlib init");
  std::string out; out.resize(std::max<size_t>(in_size * 4, 1 << 16));
  zs.next_in = (Bytef*)in; zs.avail_in = (uInt)in_size;
  size_t have = 0;
  while (true) {
    zs.next_out = (Bytef*)&out[have]; zs.avail_out = (uInt)(out.size() - have);
    int rc = inflate(&zs, Z_NO_FLUSH);
    have = out.size() - zs.avail_out;
    if (rc == Z_STREAM_END) break;
    if (rc != Z_OK) { inflateEnd(&zs); throw std::runtime_error("avro: inflate failed"); }
    if (zs.avail_out == 0) out.resize(out.size() * 2);
  }
  inflateEnd(&zs);
  out.resize(have);
  return out;
}
std::string inflate_raw(const std::string& in) { return inflate_raw(in.data(), in.size()); }
std::string deflate_raw(const std::string& in, int level) {
  z_stream zs; memset(&zs, 0, sizeof(zs));
  if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("zlib init");
  std::string out; out.resize(deflateBound(&zs, (uLong)in.size()));
  zs.next_in = (Bytef*)in.data(); zs.avail_in = (uInt)in.size();
  zs.next_out = (Bytef*)&out[0]; zs.avail_out = (uInt)out.size();
  if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); throw std::runtime_error("avro: deflate failed"); }
  out.resize(out.size() - zs.avail_out);
  deflateEnd(&zs);
  return out;
}
// container header: magic, metadata map (schema, codec), sync marker; returns the offset of the first block
size_t parse_header(const std::string& data, const std::string& path, std::string& schema_json, std::string& codec, std::string& sync) {
  if (data.size() < 4 || data.compare(0, 4, "Obj\x01")) throw std::runtime_error(path + ": not an avro container file");
  size_t pos = 4;
  codec = "null";
  while (true) {
    int64_t n = rd_long(data, pos);
    if (n == 0) break;
    if (n < 0) { n = -n; rd_long(data, pos); }
    for (int64_t k = 0; k < n; k++) {
      int64_t kl = rd_long(data, pos); need(data, pos, kl); std::string key(data, pos, (size_t)kl); pos += (size_t)kl;
      int64_t vl = rd_long(data, pos); need(data, pos, vl); std::string val(data, pos, (size_t)vl); pos += (size_t)vl;
      if (key == "avro.schema") schema_json = val;
      if (key == "avro.codec") codec = val;
    }
  }
  if (codec != "null" && codec != "deflate") throw std::runtime_error("avro: unsupported codec " + codec);
  need(data, pos, 16);
  sync.assign(data, pos, 16); pos += 16;
  return pos;
}
std::string read_whole_file(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  std::string data;
  struct stat st;
  if (fstat(fileno(f), &st) == 0 && st.st_size > 0) data.reserve((size_t)st.st_size);
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.append(buf, n);
  fclose(f);
  return data;
}
}  // namespace

static std::atomic<int> g_deflate_level{1};
int default_deflate_level() { return g_deflate_level.load(); }
void set_default_deflate_level(int level) { g_deflate_level.store(std::max(0, std::min(level, 9))); }
static std::atomic<int> g_host_threads{0};   // 0 = not decided yet
void set_host_threads(int n) { g_host_threads.store(std::max(0, std::min(n, 64))); }
int host_threads() {
  int n = g_host_threads.load();
  if (n <= 0) {
    const char* e = getenv("MLEASE_HOST_THREADS");
    int v = e ? atoi(e) : 0;
    if (v <= 0) {   // the CPUs this process may run on (a container's share), not the machine's
      cpu_set_t set;
      CPU_ZERO(&set);
      v = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
    }
    n = std::max(1, std::min(v, 64));
    g_host_threads.store(n);
  }
  return n;
}

AvroReader::AvroReader(const std::string& path) {
  data_ = read_whole_file(path);
  pos_ = parse_header(data_, path, schema_json_, codec_, sync_);
  schema_ = schema_parse(schema_json_);
}
bool AvroReader::load_block() {
  if (pos_ >= data_.size()) return false;
  remaining_ = rd_long(data_, pos_);
  int64_t bytes = rd_long(data_, pos_);
  if (remaining_ < 0) throw std::runtime_error("avro: negative record count in block header");
  need(data_, pos_, bytes);
  std::string raw(data_, pos_, (size_t)bytes); pos_ += (size_t)bytes;
  need(data_, pos_, 16);
  if (data_.compare(pos_, 16, sync_)) throw std::runtime_error("avro: sync marker mismatch");
  pos_ += 16;
  block_ = codec_ == "deflate" ? inflate_raw(raw) : raw;
  bpos_ = 0; blocks_++;
  return true;
}
bool AvroReader::next(Value& out) {
  while (remaining_ == 0) if (!load_block()) return false;
  decode(block_, bpos_, *schema_, out);
  remaining_--;
  return true;
}

AvroFile::AvroFile(const std::string& path) {
  data_ = read_whole_file(path);
  std::string sync;
  size_t pos = parse_header(data_, path, schema_json_, codec_, sync);
  schema_ = schema_parse(schema_json_);
  int64_t before = 0;
  while (pos < data_.size()) {
    int64_t count = rd_long(data_, pos);
    int64_t bytes = rd_long(data_, pos);
    if (count < 0) throw std::runtime_error("avro: negative record count in block header");
    need(data_, pos, bytes);
    const size_t off = pos;
    pos += (size_t)bytes;
    need(data_, pos, 16);
    if (data_.compare(pos, 16, sync)) throw std::runtime_error("avro: sync marker mismatch");
    pos += 16;
    blocks_.push_back(Blk{off, (size_t)bytes, count, before});
    before += count;
  }
}
std::string AvroFile::block_data(size_t b) const {
  const Blk& k = blocks_[b];
  return codec_ == "deflate" ? inflate_raw(data_.data() + k.off, k.bytes) : std::string(data_, k.off, k.bytes);
}

struct AvroWriter::Pending {
  std::future<std::string> payload;
  int64_t count = 0;
};

AvroWriter::AvroWriter(const std::string& path, const std::string& schema_json, const std::string& codec, int level)
    : path_(path), codec_(codec), level_(level < 0 ? default_deflate_level() : level) {
  schema_ = schema_parse(schema_json);
  std::mt19937_64 rng(0x6d6c65617365ULL ^ std::hash<std::string>()(path));
  sync_.resize(16);
  for (int k = 0; k < 16; k++) sync_[k] = (char)(rng() & 0xFF);
  out_ = "Obj\x01";
  wr_long(out_, 2);
  auto kv = [&](const std::string& k, const std::string& v) { wr_long(out_, (int64_t)k.size()); out_ += k; wr_long(out_, (int64_t)v.size()); out_ += v; };
  kv("avro.schema", schema_json); kv("avro.codec", codec_);
  wr_long(out_, 0);
  out_ += sync_;
}
AvroWriter::~AvroWriter() { try { close(); } catch (...) {} }
void AvroWriter::append(const Value& v) {
  encode(buf_, *schema_, v);
  if (++count_ >= 4096 || buf_.size() > (1 << 20)) flush_block();
}
void AvroWriter::append_encoded(const char* bytes, size_t nbytes, int64_t nrecords) {
  if (nrecords <= 0) return;
  buf_.append(bytes, nbytes);
  count_ += nrecords;
  if (count_ >= 4096 || buf_.size() > (1 << 20)) flush_block();
}
// appends finished blocks to the file image in order until at most `keep` are still being compressed
void AvroWriter::drain(size_t keep) {
  size_t done = 0;
  while (pending_.size() - done > keep) {
    Pending& p = *pending_[done];
    std::string payload = p.payload.get();
    wr_long(out_, p.count); wr_long(out_, (int64_t)payload.size()); out_ += payload; out_ += sync_;
    done++;
  }
  pending_.erase(pending_.begin(), pending_.begin() + (long)done);
}
void AvroWriter::flush_block() {
  if (count_ == 0) return;
  auto p = std::make_unique<Pending>();
  p->count = count_;
  if (codec_ == "deflate") {
    const int level = level_;
    auto raw = std::make_shared<std::string>(std::move(buf_));
    p->payload = std::async(std::launch::async, [raw, level]() { return deflate_raw(*raw, level); });
  } else {
    std::promise<std::string> pr;
    pr.set_value(std::move(buf_));
    p->payload = pr.get_future();
  }
  pending_.push_back(std::move(p));
  buf_ = std::string(); count_ = 0;
  drain((size_t)host_threads());
}
void AvroWriter::close() {
  if (closed_) return;
  flush_block();
  drain(0);
  size_t sl = path_.rfind('/');
  if (sl != std::string::npos) make_dirs(path_.substr(0, sl));
  std::ofstream f(path_, std::ios::binary | std::ios::trunc);
  if (!f) throw std::runtime_error("cannot write " + path_);
  f.write(out_.data(), (std::streamsize)out_.size());
  closed_ = true;
}

// ------------------------------------------------------------------------------------------ filesystem helpers
bool path_exists(const std::string& p) { struct stat st; return !p.empty() && stat(p.c_str(), &st) == 0; }
void make_dirs(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); i++) {
    if (i == path.size() || path[i] == '/') { if (!cur.empty() && !path_exists(cur)) mkdir(cur.c_str(), 0755); }
    if (i < path.size()) cur.push_back(path[i]);
  }
}
std::vector<std::string> list_avro_files(const std::string& path) {
  std::vector<std::string> out;
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return out;
  if (!S_ISDIR(st.st_mode)) { out.push_back(path); return out; }
  DIR* d = opendir(path.c_str());
  if (!d) return out;
  while (dirent* e = readdir(d)) {
    std::string n = e->d_name;
    if (n.empty() || n[0] == '.' || n[0] == '_') continue;
    std::string full = path + "/" + n;
    if (stat(full.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) { auto sub = list_avro_files(full); out.insert(out.end(), sub.begin(), sub.end()); }
    else out.push_back(full);
  }
  closedir(d);
  std::sort(out.begin(), out.end());
  return out;
}
void remove_tree(const std::string& path) {
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return;
  if (S_ISDIR(st.st_mode)) {
    DIR* d = opendir(path.c_str());
    if (d) {
      while (dirent* e = readdir(d)) { std::string n = e->d_name; if (n == "." || n == "..") continue; remove_tree(path + "/" + n); }
      closedir(d);
    }
    rmdir(path.c_str());
  } else unlink(path.c_str());
}

}  // namespace mlease_host
