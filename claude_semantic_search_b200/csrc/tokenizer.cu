// tokenizer.cu -- native batch WordPiece tokenizer (host code; SURVEY 8(f) row 2: at >= 10k chunks/s per
// GPU the Python tokenizer inside SentenceTransformer.encode is the bottleneck of the indexing path).
//
// What MPNetTokenizer / BertTokenizer do for all-mpnet-base-v2 (transformers tokenization_bert.py
// BasicTokenizer + WordpieceTokenizer): clean, lower-case, strip accents, split on whitespace and
// punctuation, greedy longest-match WordPiece with the "##" continuation prefix, [UNK] for words
// longer than 100 characters or without a match, <s> ... </s> framing, truncation.
//
// Scope: texts made of printable ASCII, space, \t, \n, \r -- for those every Unicode rule of the
// reference (NFD accent stripping, category-P punctuation, CJK spacing, control-character
// removal) is the identity or a fixed ASCII table, so the native result is the reference's by
// construction.  Any other byte flags the text (needs_fallback[i] = 1, zero tokens emitted) and
// the host tokenises it with the Python implementation: never a silent approximation.
#include "css_common.cuh"

#include <algorithm>
#include <cstdio>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

using namespace css;

struct css_tokenizer {
  std::unordered_map<std::string, int32_t> vocab;
  int lower = 1;
  int32_t bos = 0, eos = 2, unk = 3;
  size_t max_piece = 0;   // longest vocabulary entry (bytes), bounds the longest-match search
};

namespace {

inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
inline bool is_punct(unsigned char c) {
  return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126);
}
// printable ASCII or one of the four whitespace characters the reference keeps as separators
inline bool in_scope(unsigned char c) { return (c >= 32 && c <= 126) || c == '\t' || c == '\n' || c == '\r'; }

// greedy longest-match WordPiece of one word (already lower-cased); appends to `out`
void wordpiece(const css_tokenizer& t, const char* w, size_t len, std::string& scratch, std::vector<int32_t>& out) {
  if (len > 100) {
    out.push_back(t.unk);
    return;
  }
  const size_t first = out.size();
  size_t start = 0;
  while (start < len) {
    size_t end = len;
    int32_t id = -1;
    while (start < end) {
      const size_t plen = end - start + (start > 0 ? 2 : 0);
      if (plen <= t.max_piece) {
        scratch.clear();
        if (start > 0) scratch.append("##");
        scratch.append(w + start, end - start);
        auto it = t.vocab.find(scratch);
        if (it != t.vocab.end()) {
          id = it->second;
          break;
        }
      }
      --end;
    }
    if (id < 0) {   // no piece matches: the whole word is unknown
      out.resize(first);
      out.push_back(t.unk);
      return;
    }
    out.push_back(id);
    start = end;
  }
}

// returns false when the text is out of scope (non-ASCII / control bytes)
bool encode_one(const css_tokenizer& t, const char* text, int64_t len, int max_len, std::string& word,
                std::string& scratch, std::vector<int32_t>& out) {
  for (int64_t i = 0; i < len; ++i)
    if (!in_scope((unsigned char)text[i])) return false;
  const size_t budget = (size_t)std::max(max_len - 2, 0);
  const size_t base = out.size();
  out.push_back(t.bos);
  word.clear();
  auto flush = [&]() {
    if (!word.empty()) {
      wordpiece(t, word.data(), word.size(), scratch, out);
      word.clear();
    }
  };
  for (int64_t i = 0; i < len && out.size() - base - 1 < budget; ++i) {
    unsigned char c = (unsigned char)text[i];
    if (is_space(c)) {
      flush();
    } else if (is_punct(c)) {
      flush();
      if (out.size() - base - 1 >= budget) break;
      const char p = (char)c;
      wordpiece(t, &p, 1, scratch, out);
    } else {
      word.push_back(t.lower && c >= 'A' && c <= 'Z' ? (char)(c + 32) : (char)c);
    }
  }
  if (out.size() - base - 1 < budget) flush();
  if (out.size() - base - 1 > budget) out.resize(base + 1 + budget);
  out.push_back(t.eos);
  return true;
}

}  // namespace

extern "C" {

int css_tokenizer_create(const char* vocab_path, int do_lower_case, css_tokenizer** out) {
  CSS_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  CSS_REQUIRE(vocab_path != nullptr, "vocab_path is NULL");
  FILE* f = fopen(vocab_path, "rb");
  if (!f) {
    set_error("cannot open vocabulary file %s", vocab_path);
    return CSS_ERR_IO;
  }
  css_tokenizer* t = new (std::nothrow) css_tokenizer();
  if (!t) {
    fclose(f);
    set_error("out of host memory");
    return CSS_ERR_OOM;
  }
  t->lower = do_lower_case ? 1 : 0;
  std::string line;
  int32_t idx = 0;
  int ch;
  auto commit = [&]() {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    t->vocab[line] = idx++;   // a duplicate keeps the LAST index, like the Python dict the reference builds
    t->max_piece = std::max(t->max_piece, line.size());
    line.clear();
  };
  bool pending = false;
  while ((ch = fgetc(f)) != EOF) {
    if (ch == '\n') {
      commit();
      pending = false;
    } else {
      line.push_back((char)ch);
      pending = true;
    }
  }
  if (pending) commit();
  fclose(f);
  auto find = [&](const char* s, int32_t dflt) {
    auto it = t->vocab.find(s);
    return it == t->vocab.end() ? dflt : it->second;
  };
  t->bos = find("<s>", 0);
  t->eos = find("</s>", 2);
  t->unk = find("[UNK]", find("<unk>", 3));
  *out = t;
  return CSS_OK;
}

int css_tokenizer_destroy(css_tokenizer* t) {
  delete t;
  return CSS_OK;
}

int css_tokenizer_vocab_size(const css_tokenizer* t) { return t ? (int)t->vocab.size() : CSS_ERR_INVALID; }

int css_tokenizer_encode_batch(css_tokenizer* t, const char* const* texts, const int64_t* lens, int32_t n,
                               int32_t max_len, int32_t* ids_out, int32_t* cu_seqlens_out, uint8_t* needs_fallback,
                               int32_t n_threads) {
  CSS_REQUIRE(t != nullptr, "tokenizer is NULL");
  CSS_REQUIRE(n >= 0 && max_len >= 2, "bad n / max_len");
  CSS_REQUIRE(cu_seqlens_out != nullptr, "cu_seqlens_out is NULL");
  cu_seqlens_out[0] = 0;
  if (n == 0) return CSS_OK;
  CSS_REQUIRE(texts && lens && ids_out && needs_fallback, "NULL buffer");
  int nt = n_threads > 0 ? n_threads : (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
  nt = std::max(1, std::min(nt, (int)((n + 63) / 64)));
  std::vector<std::vector<int32_t>> part(nt);
  std::vector<std::vector<int32_t>> part_len(nt);
  auto work = [&](int w) {
    const int32_t i0 = (int32_t)((int64_t)n * w / nt), i1 = (int32_t)((int64_t)n * (w + 1) / nt);
    std::string word, scratch;
    std::vector<int32_t>& ids = part[w];
    std::vector<int32_t>& ln = part_len[w];
    ids.reserve((size_t)(i1 - i0) * 64);
    ln.reserve(i1 - i0);
    for (int32_t i = i0; i < i1; ++i) {
      const size_t before = ids.size();
      const bool ok = encode_one(*t, texts[i], lens[i], max_len, word, scratch, ids);
      if (!ok) ids.resize(before);
      needs_fallback[i] = ok ? 0 : 1;
      ln.push_back((int32_t)(ids.size() - before));
    }
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int w = 0; w < nt; ++w) th.emplace_back(work, w);
    for (auto& x : th) x.join();
  }
  int64_t pos = 0;
  int32_t i = 0;
  for (int w = 0; w < nt; ++w) {
    std::copy(part[w].begin(), part[w].end(), ids_out + pos);
    int64_t p = pos;
    for (int32_t l : part_len[w]) {
      p += l;
      cu_seqlens_out[++i] = (int32_t)p;
    }
    pos += (int64_t)part[w].size();
  }
  return CSS_OK;
}

}  // extern "C"
