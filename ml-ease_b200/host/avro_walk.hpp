// avro_walk.hpp -- allocation-free, block-parallel reading of Avro records for the hot ingest paths of the job layer.
//
// The generic decoder of avro_io builds a Value tree per record (two std::strings and a vector per node): fine for models and
// lambda maps, far too slow for the data files, which are where the reference spends its wall clock (every AdmmReducer re-reads
// its partition through LibLinearDataset.addInstanceAvro in every iteration, llf/LibLinearDataset.java:413-484).  Here a schema
// is compiled once into a Plan tree in which the handful of leaves a job reads (key, response, weight, offset, the
// name / term / value of a feature) carry a slot number; walking the plan over a block's bytes stores numbers and string VIEWS
// (pointers into the block) into the slots and reports every element of a tagged array to the caller's sink.  Nothing is
// allocated per record.  Blocks of a container file are independent, so they are decoded on all host threads; the results are
// merged in block order, which keeps every order-dependent result (first-seen feature ids, record order) identical to a
// sequential read.  Anything unusual in a schema (a recursive type, a non-scalar where a job reads a scalar) makes the caller
// fall back to the generic decoder, which also serves as the reference implementation in the tests.
#pragma once
#include <atomic>
#include <cstring>
#include <exception>
#include <functional>
#include <thread>

#include "avro_io.hpp"

namespace mlease_host {

struct Slot {
  enum Kind : uint8_t { Null, Bool, Int, Long, Float, Double, Str, Unset } kind = Null;
  int64_t i = 0;            // boolean / int / long / enum index
  double d = 0;             // float / double
  const char* p = nullptr;  // string / bytes / fixed: view into the block
  size_t n = 0;
  bool is_null() const { return kind == Null; }
  double num() const { return (kind == Float || kind == Double) ? d : (double)i; }   // num_of() of the generic path
};

struct Plan {
  Schema::Type type = Schema::Null;
  int tag = -1;             // scalar leaf: slot to capture into; array: event id (begin_item / item calls of the sink)
  int fixed_size = 0;
  std::vector<Plan> kids;   // record: fields; union: branches; array / map: the item plan
};

inline bool plan_is_scalar(Schema::Type t) {
  return t == Schema::Null || t == Schema::Boolean || t == Schema::Int || t == Schema::Long || t == Schema::Float || t == Schema::Double ||
         t == Schema::String || t == Schema::Bytes || t == Schema::Enum || t == Schema::Fixed;
}

// Plan tree of a schema.  Throws on recursive (self-referencing) types.
inline Plan plan_build(const Schema& s, int depth = 0) {
  if (depth > 24) throw std::runtime_error("avro: schema too deep / recursive for the plan walker");
  Plan p;
  p.type = s.type;
  p.fixed_size = s.fixed_size;
  switch (s.type) {
    case Schema::Record: for (auto& f : s.fields) p.kids.push_back(plan_build(*f.second, depth + 1)); break;
    case Schema::Union: for (auto& b : s.branches) p.kids.push_back(plan_build(*b, depth + 1)); break;
    case Schema::Array: case Schema::Map: p.kids.push_back(plan_build(*s.items, depth + 1)); break;
    default: break;
  }
  return p;
}

// The record behind unions (first non-null branch at every level, like rec_schema() of the job layer): plan node + schema node.
inline bool plan_resolve(Plan*& p, const Schema*& s, Schema::Type want) {
  while (s->type == Schema::Union) {
    int nx = -1;
    for (size_t b = 0; b < s->branches.size(); b++) if (s->branches[b]->type != Schema::Null) { nx = (int)b; break; }
    if (nx < 0) return false;
    p = &p->kids[nx];
    s = s->branches[nx].get();
  }
  return s->type == want;
}
// Tags every scalar leaf reachable through unions from this node with `slot`; false if a non-null branch is not a scalar.
inline bool plan_tag_scalar(Plan& p, int slot) {
  if (p.type == Schema::Union) {
    for (auto& k : p.kids) if (!plan_tag_scalar(k, slot)) return false;
    return true;
  }
  if (!plan_is_scalar(p.type)) return false;
  p.tag = slot;
  return true;
}
// Tags field `name` of record (rp, rs) as a scalar slot.  Returns 0 if the record has no such field (the slot then keeps its
// initial Null), 1 if tagged, -1 if the field is not a scalar (caller falls back to the generic path).
inline int plan_tag_field(Plan& rp, const Schema& rs, const std::string& name, int slot) {
  const int i = rs.field_index(name);
  if (i < 0) return 0;
  return plan_tag_scalar(rp.kids[i], slot) ? 1 : -1;
}

namespace walk_detail {
[[noreturn]] inline void overrun() { throw std::runtime_error("avro: truncated or corrupt block (value runs past the end of the data)"); }
inline int64_t rd_long(const uint8_t*& p, const uint8_t* e) {
  uint64_t acc = 0;
  int sh = 0;
  while (true) {
    if (p >= e) throw std::runtime_error("avro: truncated varint");
    const uint8_t c = *p++;
    if (sh > 63) throw std::runtime_error("avro: varint longer than 10 bytes");
    acc |= (uint64_t)(c & 0x7F) << sh;
    if (!(c & 0x80)) break;
    sh += 7;
  }
  return (int64_t)(acc >> 1) ^ -(int64_t)(acc & 1);
}
inline void need(const uint8_t* p, const uint8_t* e, int64_t n) {
  if (n < 0 || (uint64_t)n > (uint64_t)(e - p)) overrun();
}
}  // namespace walk_detail

// Sink: void begin_item(int event, Slot* slots); void item(int event, Slot* slots);
template <class Sink>
inline void plan_walk(const Plan& pl, const uint8_t*& p, const uint8_t* e, Slot* slots, Sink& sink) {
  using namespace walk_detail;
  switch (pl.type) {
    case Schema::Null:
      if (pl.tag >= 0) slots[pl.tag].kind = Slot::Null;
      break;
    case Schema::Boolean:
      need(p, e, 1);
      if (pl.tag >= 0) { Slot& s = slots[pl.tag]; s.kind = Slot::Bool; s.i = *p != 0; }
      p++;
      break;
    case Schema::Int: case Schema::Long: case Schema::Enum: {
      const int64_t v = rd_long(p, e);
      if (pl.tag >= 0) { Slot& s = slots[pl.tag]; s.kind = pl.type == Schema::Int ? Slot::Int : Slot::Long; s.i = v; }
      break;
    }
    case Schema::Float: {
      need(p, e, 4);
      if (pl.tag >= 0) { float f; std::memcpy(&f, p, 4); Slot& s = slots[pl.tag]; s.kind = Slot::Float; s.d = f; }
      p += 4;
      break;
    }
    case Schema::Double: {
      need(p, e, 8);
      if (pl.tag >= 0) { double d; std::memcpy(&d, p, 8); Slot& s = slots[pl.tag]; s.kind = Slot::Double; s.d = d; }
      p += 8;
      break;
    }
    case Schema::String: case Schema::Bytes: {
      const int64_t n = rd_long(p, e);
      need(p, e, n);
      if (pl.tag >= 0) { Slot& s = slots[pl.tag]; s.kind = Slot::Str; s.p = reinterpret_cast<const char*>(p); s.n = (size_t)n; }
      p += n;
      break;
    }
    case Schema::Fixed:
      need(p, e, pl.fixed_size);
      if (pl.tag >= 0) { Slot& s = slots[pl.tag]; s.kind = Slot::Str; s.p = reinterpret_cast<const char*>(p); s.n = (size_t)pl.fixed_size; }
      p += pl.fixed_size;
      break;
    case Schema::Union: {
      const int64_t br = rd_long(p, e);
      if (br < 0 || br >= (int64_t)pl.kids.size()) throw std::runtime_error("avro: bad union branch");
      plan_walk(pl.kids[(size_t)br], p, e, slots, sink);
      break;
    }
    case Schema::Record:
      for (auto& k : pl.kids) plan_walk(k, p, e, slots, sink);
      break;
    case Schema::Array:
      while (true) {
        int64_t n = rd_long(p, e);
        if (n == 0) break;
        if (n < 0) { n = -n; rd_long(p, e); }
        for (int64_t k = 0; k < n; k++) {
          if (pl.tag >= 0) sink.begin_item(pl.tag, slots);
          plan_walk(pl.kids[0], p, e, slots, sink);
          if (pl.tag >= 0) sink.item(pl.tag, slots);
        }
      }
      break;
    case Schema::Map:
      while (true) {
        int64_t n = rd_long(p, e);
        if (n == 0) break;
        if (n < 0) { n = -n; rd_long(p, e); }
        for (int64_t k = 0; k < n; k++) {
          const int64_t l = rd_long(p, e);
          need(p, e, l);
          p += l;
          plan_walk(pl.kids[0], p, e, slots, sink);
        }
      }
      break;
  }
}

// Strings -> dense ids in insertion order (open addressing over an arena; lookups take a view, nothing is allocated per call).
class StrTable {
 public:
  StrTable() { slot_.assign(1024, -1); }
  static uint64_t hash_of(const char* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < n; i++) { h ^= (uint8_t)p[i]; h *= 0x100000001b3ULL; }
    return h ^ (h >> 29);
  }
  int find_or_add(const char* p, size_t n) {
    const uint64_t h = hash_of(p, n);
    size_t m = slot_.size() - 1, i = (size_t)h & m;
    while (true) {
      const int id = slot_[i];
      if (id < 0) break;
      if (hash_[id] == h && len_[id] == n && std::memcmp(arena_.data() + off_[id], p, n) == 0) return id;
      i = (i + 1) & m;
    }
    const int id = (int)off_.size();
    off_.push_back(arena_.size()); len_.push_back(n); hash_.push_back(h);
    arena_.append(p, n);
    slot_[i] = id;
    if (off_.size() * 2 > slot_.size()) grow();
    return id;
  }
  size_t size() const { return off_.size(); }
  const char* data(int id) const { return arena_.data() + off_[id]; }
  size_t length(int id) const { return len_[id]; }
 private:
  void grow() {
    std::vector<int32_t> ns(slot_.size() * 2, -1);
    const size_t m = ns.size() - 1;
    for (size_t id = 0; id < off_.size(); id++) {
      size_t i = (size_t)hash_[id] & m;
      while (ns[i] >= 0) i = (i + 1) & m;
      ns[i] = (int32_t)id;
    }
    slot_.swap(ns);
  }
  std::vector<int32_t> slot_;
  std::vector<uint64_t> hash_;
  std::vector<size_t> off_, len_;
  std::string arena_;
};

// fn(block) for every block index on up to `threads` threads (dynamic order).  If any call throws, the exception of the LOWEST
// block index is rethrown after all workers have stopped: the error a sequential reader would have hit first.
inline void parallel_blocks(size_t nblocks, int threads, const std::function<void(size_t)>& fn) {
  if (nblocks == 0) return;
  const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), nblocks));
  std::vector<std::exception_ptr> errs(nblocks);
  std::atomic<size_t> next{0};
  std::atomic<size_t> first_bad{nblocks};
  auto work = [&]() {
    while (true) {
      const size_t b = next.fetch_add(1);
      if (b >= nblocks || b > first_bad.load()) break;   // blocks after a failed one are not needed
      try { fn(b); }
      catch (...) {
        errs[b] = std::current_exception();
        size_t cur = first_bad.load();
        while (b < cur && !first_bad.compare_exchange_weak(cur, b)) {}
      }
    }
  };
  if (nt == 1) work();
  else {
    std::vector<std::thread> ts;
    for (int t = 0; t < nt; t++) ts.emplace_back(work);
    for (auto& t : ts) t.join();
  }
  for (size_t b = 0; b < nblocks; b++) if (errs[b]) std::rethrow_exception(errs[b]);
}

}  // namespace mlease_host
