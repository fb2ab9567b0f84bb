#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace vox {

class Tokenizer {
  public:
    static constexpr uint32_t kTextTokenOffset = 1000;  // tokenizer/mod.rs:66
    static Tokenizer *from_json(const char *json, size_t len);
    static Tokenizer *from_file(const std::string &path);
    std::string decode(const uint32_t *ids, size_t n) const;
    bool decode_token(uint32_t id, std::string *out) const;
    size_t vocab_size() const { return vocab_size_; }

  private:
    std::vector<std::string> vocab_bytes_;
    std::vector<uint8_t> has_bytes_;
    std::map<uint32_t, std::string> special_;
    size_t vocab_size_ = 0;
};

}  // namespace vox
