// gguf.h -- GGUF v2/v3 reader over a file (mmap) or a list of borrowed in-memory shards.
// Mirrors GgufReader / ShardedCursor (reference src/gguf/reader.rs:88-314): header, metadata
// KVs, tensor index, 32-byte aligned data section, random access by tensor name.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace vox {

struct GgufTensorInfo {
    std::string name;
    std::vector<uint64_t> dims;  // GGUF order (innermost first), as stored
    uint32_t dtype = 0;          // 0 F32, 1 F16, 2 Q4_0
    uint64_t offset = 0;         // relative to the data section
    uint64_t num_elements() const;
    uint64_t byte_size() const;
    // PyTorch-order shape (reverse_gguf_dims, loader.rs:497-499)
    std::vector<int64_t> shape() const;
};

class Gguf {
  public:
    static Gguf *open_file(const std::string &path);
    static Gguf *open_shards(const void *const *bufs, const size_t *lens, size_t n);
    ~Gguf();

    uint32_t version() const { return version_; }
    uint64_t tensor_count() const { return tensor_count_; }
    const GgufTensorInfo *find(const std::string &name) const;
    const std::vector<std::string> &names() const { return names_; }
    // copy `byte_size` bytes of tensor data to dst
    void read_tensor(const GgufTensorInfo &t, void *dst) const;
    // u32 metadata value (our synthetic models' optional `voxtral.*` keys); false if absent
    bool kv_u32(const std::string &key, uint32_t *out) const;

  private:
    Gguf() = default;
    void parse();
    void read_at(uint64_t pos, void *dst, size_t n) const;  // ShardedCursor read+seek
    uint64_t total_len_ = 0;
    // shards: base pointers + cumulative ends (ShardedCursor::ends)
    std::vector<const uint8_t *> shard_ptr_;
    std::vector<uint64_t> shard_end_;
    void *map_base_ = nullptr;  // mmap of a file, if any
    size_t map_len_ = 0;
    uint32_t version_ = 0;
    uint64_t tensor_count_ = 0;
    uint64_t data_offset_ = 0;
    std::map<std::string, GgufTensorInfo> tensors_;
    std::vector<std::string> names_;
    std::map<std::string, uint32_t> kv_u32_;
};

}  // namespace vox
