// avro_io.hpp -- minimal Avro object-container codec for the job layer (null + deflate codecs, zig-zag varints,
// schema-driven generic decode/encode incl. Pig-style ["null", T] unions).  Replaces the reference's use of
// avro 1.7.6 + AvroHdfsFileReader/Writer (com/linkedin/mapred/Avro*.java) for local files.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace mlease_host {

// ---------------------------------------------------------------- tiny JSON (schemas only)
struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;
  const Json* get(const std::string& k) const {
    for (auto& kv : obj) if (kv.first == k) return &kv.second;
    return nullptr;
  }
};
Json json_parse(const std::string& text);
std::string json_dump(const Json& j);

// ---------------------------------------------------------------- schema + generic values
struct Schema;
using SchemaP = std::shared_ptr<Schema>;
struct Schema {
  enum Type { Null, Boolean, Int, Long, Float, Double, String, Bytes, Record, Array, Union, Map, Enum, Fixed } type = Null;
  std::string name;                                   // records / enums / fixed
  std::vector<std::pair<std::string, SchemaP>> fields;  // records
  SchemaP items;                                      // arrays / map values
  std::vector<SchemaP> branches;                      // unions
  std::vector<std::string> symbols;                   // enums
  int fixed_size = 0;
  int field_index(const std::string& n) const {
    for (size_t i = 0; i < fields.size(); i++) if (fields[i].first == n) return (int)i;
    return -1;
  }
};
SchemaP schema_from_json(const Json& j, std::map<std::string, SchemaP>& named);
SchemaP schema_parse(const std::string& json_text);
Json schema_to_json(const SchemaP& s, std::map<std::string, bool>& emitted);
SchemaP schema_remove_union(const SchemaP& s);   // utils/Util.java:377-417 (Util.removeUnion)

struct Value {
  Schema::Type type = Schema::Null;
  int64_t i = 0;       // boolean / int / long / enum index
  double d = 0;        // float / double
  std::string s;       // string / bytes / fixed
  std::string map_key; // key of this value when it is an entry of a map
  std::vector<Value> items;   // array elements, record fields (by index), map values (map_key = key of each item)
  bool is_null() const { return type == Schema::Null; }
  static Value null() { return Value(); }
  static Value of_int(int64_t v) { Value x; x.type = Schema::Int; x.i = v; return x; }
  static Value of_float(float v) { Value x; x.type = Schema::Float; x.d = v; return x; }
  static Value of_double(double v) { Value x; x.type = Schema::Double; x.d = v; return x; }
  static Value of_string(const std::string& v) { Value x; x.type = Schema::String; x.s = v; return x; }
};

// ---------------------------------------------------------------- container files
class AvroReader {
 public:
  explicit AvroReader(const std::string& path);
  const SchemaP& schema() const { return schema_; }
  const std::string& schema_json() const { return schema_json_; }
  bool next(Value& out);          // false at end of file
  int64_t blocks_read() const { return blocks_; }
 private:
  bool load_block();
  std::string data_;
  size_t pos_ = 0;
  std::string block_;
  size_t bpos_ = 0;
  int64_t remaining_ = 0, blocks_ = 0;
  std::string sync_, codec_, schema_json_;
  SchemaP schema_;
};

class AvroWriter {
 public:
  // codec: "null" or "deflate" (the reference's jobs write deflate, com/linkedin/mapred/AbstractAvroJob.java:253).  Blocks are
  // compressed on background threads (each block is an independent raw-deflate stream, so the file is the same as a serial
  // writer's); level = zlib level of the deflate codec.
  AvroWriter(const std::string& path, const std::string& schema_json, const std::string& codec = "deflate", int level = -1);   // -1: default_deflate_level()
  ~AvroWriter();
  void append(const Value& v);
  // records already in Avro binary form for this writer's schema (block-parallel producers encode on their own threads)
  void append_encoded(const char* bytes, size_t nbytes, int64_t nrecords);
  void close();
  const SchemaP& schema() const { return schema_; }
 private:
  void flush_block();
  void drain(size_t keep);
  struct Pending;
  std::string path_, codec_, buf_, out_;
  SchemaP schema_;
  std::string sync_;
  int64_t count_ = 0;
  int level_ = 6;
  bool closed_ = false;
  std::vector<std::unique_ptr<Pending>> pending_;   // blocks being compressed, in file order
};

// Random access to the blocks of a container file: the whole file is read, the header parsed and every block header / sync
// marker checked up front (a truncated or corrupt file is an error here, never an out-of-bounds read later); block payloads
// are inflated on demand by whichever thread asks (const, thread-safe), which is what the block-parallel readers of the job
// layer build on.
class AvroFile {
 public:
  explicit AvroFile(const std::string& path);
  const SchemaP& schema() const { return schema_; }
  const std::string& schema_json() const { return schema_json_; }
  size_t num_blocks() const { return blocks_.size(); }
  int64_t block_records(size_t b) const { return blocks_[b].count; }
  int64_t records_before(size_t b) const { return blocks_[b].before; }
  int64_t num_records() const { return blocks_.empty() ? 0 : blocks_.back().before + blocks_.back().count; }
  std::string block_data(size_t b) const;   // decompressed payload of block b
 private:
  struct Blk { size_t off, bytes; int64_t count, before; };
  std::string data_, codec_, schema_json_;
  SchemaP schema_;
  std::vector<Blk> blocks_;
};

// zlib level of the files the job layer writes.  Default 1: on model-like data (short names + floats) level 1 compresses to the
// same size as 6 or 9 (0.65 vs 0.63 of the raw bytes) at 4x the speed; the reference asks for 9 on HDFS
// (com/linkedin/mapred/AbstractAvroJob.java:253), which only changes the bytes of the file, not the records.  Job key: avro.deflate.level.
int default_deflate_level();
void set_default_deflate_level(int level);
// worker threads of the host layer: set_host_threads(n > 0), else MLEASE_HOST_THREADS, else the CPUs this process may run on (at most 64)
int host_threads();
void set_host_threads(int n);   // 0 = back to the default

// helpers
std::vector<std::string> list_avro_files(const std::string& path);   // file, or *.avro / part-* files of a directory (sorted)
void make_dirs(const std::string& path);
bool path_exists(const std::string& path);
void remove_tree(const std::string& path);

}  // namespace mlease_host
