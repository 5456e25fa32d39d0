// tokenizer.cu -- native batch WordPiece tokenizer (host code; SURVEY 8(f) row 2: at >= 10k chunks/s per
// GPU the Python tokenizer inside SentenceTransformer.encode is the bottleneck of the indexing path).
//
// What MPNetTokenizer / BertTokenizer do for all-mpnet-base-v2 (transformers tokenization_bert.py
// BasicTokenizer + WordpieceTokenizer): clean, lower-case, strip accents, split on whitespace and
// punctuation, greedy longest-match WordPiece with the "##" continuation prefix, [UNK] for words
// longer than 100 characters or without a match, <s> ... </s> framing, truncation.
//
// Scope: any valid UTF-8.  The normaliser the reference's fast tokenizer runs (Rust `tokenizers`
// BertNormalizer: control-character removal, whitespace folding, CJK padding, NFD + Mn stripping,
// per-character lower-casing) is a per-code-point map apart from NFD's canonical reordering, and so
// is BertPreTokenizer's whitespace / punctuation split; both are taken as tables generated FROM that
// pipeline (scripts/gen_unicode_tables.py -> unicode_tables.inc) and swept over every code point by
// tests/test_host_cpu.py.  Malformed UTF-8 flags the text (needs_fallback[i] = 1, zero tokens
// emitted): never a silent approximation.
#include "css_common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

using namespace css;

namespace {
#include "unicode_tables.inc"

// per-code-point entry of the two-stage table
//   bits 0-1  pre-tokeniser class of the character: 0 word, 1 whitespace, 2 punctuation
//   bits 2-3  clean-text action (both modes): 0 keep, 1 remove, 2 -> ' ', 3 pad with spaces (CJK)
//   bits 4-6  lower-case mode action: 0 none, 1 drop, 2 drop + reordering barrier, 3 Hangul NFD, 4 sequence
//   bit  7    non-starter that survives mark stripping (canonical combining class in kUniCcc)
//   bits 12+  index into kUniMapIndex for action 4
enum : uint32_t { kClsWord = 0, kClsSpace = 1, kClsPunct = 2 };
enum : uint32_t { kCleanKeep = 0, kCleanRemove = 1, kCleanSpace = 2, kCleanCjk = 3 };
enum : uint32_t { kLowNone = 0, kLowDrop = 1, kLowDropBarrier = 2, kLowHangul = 3, kLowSeq = 4 };

struct UniTable {
  std::vector<uint16_t> stage1;   // code point >> 8 -> block
  std::vector<uint32_t> blocks;   // block * 256 + (code point & 255)
  uint32_t at(uint32_t cp) const { return blocks[(size_t)stage1[cp >> 8] * 256 + (cp & 255)]; }
};

const UniTable& uni_table() {
  static UniTable tab;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<uint32_t> flat(0x110000, 0);
    auto mark = [&](const uint32_t (*r)[2], size_t n, uint32_t clear, uint32_t set) {
      for (size_t i = 0; i < n; ++i)
        for (uint32_t c = r[i][0]; c <= r[i][1]; ++c) flat[c] = (flat[c] & ~clear) | set;
    };
    mark(kUniPreSpace, kUniPreSpaceCount, 3u, kClsSpace);
    mark(kUniPrePunct, kUniPrePunctCount, 3u, kClsPunct);
    mark(kUniCleanRemove, kUniCleanRemoveCount, 3u << 2, kCleanRemove << 2);
    mark(kUniToSpace, kUniToSpaceCount, 3u << 2, kCleanSpace << 2);
    mark(kUniCjk, kUniCjkCount, 3u << 2, kCleanCjk << 2);
    mark(kUniMarkDrop, kUniMarkDropCount, 7u << 4, kLowDrop << 4);
    mark(kUniMarkDropBarrier, kUniMarkDropBarrierCount, 7u << 4, kLowDropBarrier << 4);
    for (uint32_t c = 0xAC00; c <= 0xD7A3; ++c) flat[c] |= kLowHangul << 4;
    for (size_t i = 0; i < kUniMapIndexCount; ++i)
      flat[kUniMapIndex[i][0]] = (flat[kUniMapIndex[i][0]] & 0xFFFu & ~(7u << 4)) | (kLowSeq << 4) | ((uint32_t)i << 12);
    for (size_t i = 0; i < kUniCccCount; ++i) flat[kUniCcc[i][0]] |= 1u << 7;
    tab.stage1.resize(0x1100);
    std::map<std::vector<uint32_t>, uint16_t> seen;
    for (uint32_t b = 0; b < 0x1100; ++b) {
      std::vector<uint32_t> blk(flat.begin() + (size_t)b * 256, flat.begin() + (size_t)(b + 1) * 256);
      auto it = seen.find(blk);
      if (it == seen.end()) {
        it = seen.emplace(blk, (uint16_t)seen.size()).first;
        tab.blocks.insert(tab.blocks.end(), blk.begin(), blk.end());
      }
      tab.stage1[b] = it->second;
    }
  });
  return tab;
}

inline uint32_t ccc_of(uint32_t cp) {
  for (size_t i = 0; i < kUniCccCount; ++i)
    if (kUniCcc[i][0] == cp) return kUniCcc[i][1];
  return 0;
}

}  // namespace

struct css_tokenizer {
  std::unordered_map<std::string, int32_t> vocab;
  int lower = 1;
  int32_t bos = 0, eos = 2, unk = 3;
  size_t max_piece = 0;   // longest vocabulary entry (bytes), bounds the longest-match search
  const UniTable* uni = nullptr;
  // ASCII fast path: 0 drop, 1 whitespace, 2 punctuation, 3 word character; ascii_out = the (lower-cased) byte
  uint8_t ascii_cls[128];
  char ascii_out[128];
  // special literals cut out of the raw text (longest first), and the set of their first bytes
  std::vector<std::pair<std::string, int32_t>> special;
  bool special_first[256] = {};
};

namespace {

// greedy longest-match WordPiece of one normalised word (UTF-8, `nchars` characters); appends to `out`
void wordpiece(const css_tokenizer& t, const char* w, size_t len, size_t nchars, std::string& scratch,
               std::vector<int32_t>& out) {
  if (nchars > 100) {
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
      do {   // back to the previous character boundary
        --end;
      } while (end > start && ((unsigned char)w[end] & 0xC0) == 0x80);
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

inline size_t utf8_bytes(uint32_t cp, char* b) {
  if (cp < 0x80) {
    b[0] = (char)cp;
    return 1;
  }
  if (cp < 0x800) {
    b[0] = (char)(0xC0 | (cp >> 6));
    b[1] = (char)(0x80 | (cp & 63));
    return 2;
  }
  if (cp < 0x10000) {
    b[0] = (char)(0xE0 | (cp >> 12));
    b[1] = (char)(0x80 | ((cp >> 6) & 63));
    b[2] = (char)(0x80 | (cp & 63));
    return 3;
  }
  b[0] = (char)(0xF0 | (cp >> 18));
  b[1] = (char)(0x80 | ((cp >> 12) & 63));
  b[2] = (char)(0x80 | ((cp >> 6) & 63));
  b[3] = (char)(0x80 | (cp & 63));
  return 4;
}

// strict UTF-8 decode (no overlongs, no surrogates, <= U+10FFFF); returns bytes consumed, 0 when malformed
inline int decode_utf8(const unsigned char* p, int64_t avail, uint32_t& cp) {
  const unsigned char c = p[0];
  if (c < 0x80) {
    cp = c;
    return 1;
  }
  auto cont = [&](int k) { return k < avail && (p[k] & 0xC0) == 0x80; };
  if (c >= 0xC2 && c <= 0xDF) {
    if (!cont(1)) return 0;
    cp = ((uint32_t)(c & 0x1F) << 6) | (p[1] & 0x3F);
    return 2;
  }
  if (c >= 0xE0 && c <= 0xEF) {
    if (!cont(1) || !cont(2)) return 0;
    cp = ((uint32_t)(c & 0x0F) << 12) | ((uint32_t)(p[1] & 0x3F) << 6) | (p[2] & 0x3F);
    if (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF)) return 0;
    return 3;
  }
  if (c >= 0xF0 && c <= 0xF4) {
    if (!cont(1) || !cont(2) || !cont(3)) return 0;
    cp = ((uint32_t)(c & 0x07) << 18) | ((uint32_t)(p[1] & 0x3F) << 12) | ((uint32_t)(p[2] & 0x3F) << 6) | (p[3] & 0x3F);
    if (cp < 0x10000 || cp > 0x10FFFF) return 0;
    return 4;
  }
  return 0;
}

struct Scratch {
  std::string word, piece;
  size_t word_chars = 0;
  std::vector<std::pair<uint32_t, uint32_t>> run;   // pending non-starters (code point, combining class)
};

// one text -> <s> pieces </s>; returns false when the text is malformed UTF-8
bool encode_one(const css_tokenizer& t, const char* text, int64_t len, int max_len, Scratch& s,
                std::vector<int32_t>& out) {
  const size_t budget = (size_t)std::max(max_len - 2, 0);
  const size_t base = out.size();
  const UniTable& U = *t.uni;
  out.push_back(t.bos);
  s.word.clear();
  s.word_chars = 0;
  s.run.clear();
  auto full = [&]() { return out.size() - base - 1 >= budget; };
  auto flush_word = [&]() {
    if (!s.word.empty()) {
      wordpiece(t, s.word.data(), s.word.size(), s.word_chars, s.piece, out);
      s.word.clear();
      s.word_chars = 0;
    }
  };
  // a normalised character reaches the pre-tokeniser
  auto feed = [&](uint32_t cp, uint32_t cls) {
    if (cls == kClsSpace) {
      flush_word();
    } else if (cls == kClsPunct) {
      flush_word();
      if (full()) return;
      char buf[4];
      const size_t nb = utf8_bytes(cp, buf);
      wordpiece(t, buf, nb, 1, s.piece, out);
    } else {
      char buf[4];
      s.word.append(buf, utf8_bytes(cp, buf));
      ++s.word_chars;
    }
  };
  auto flush_run = [&]() {
    if (s.run.empty()) return;
    std::stable_sort(s.run.begin(), s.run.end(), [](const auto& a, const auto& b) { return a.second < b.second; });
    for (const auto& r : s.run) feed(r.first, U.at(r.first) & 3u);
    s.run.clear();
  };
  // a character leaves the normaliser (lower-case mode runs NFD: keep non-starters sorted)
  auto emit = [&](uint32_t cp) {
    const uint32_t e = U.at(cp);
    if (t.lower && (e & (1u << 7))) {
      s.run.emplace_back(cp, ccc_of(cp));
      return;
    }
    flush_run();
    feed(cp, e & 3u);
  };
  const unsigned char* p = (const unsigned char*)text;
  int64_t i = 0;
  while (i < len && !full()) {
    const unsigned char c = p[i];
    if (t.special_first[c]) {   // an added special token ends the segment and maps straight to its id
      const std::pair<std::string, int32_t>* hit = nullptr;
      for (const auto& sp : t.special)
        if ((int64_t)sp.first.size() <= len - i && memcmp(p + i, sp.first.data(), sp.first.size()) == 0) {
          hit = &sp;
          break;
        }
      if (hit) {
        flush_run();
        flush_word();
        if (!full()) out.push_back(hit->second);
        i += (int64_t)hit->first.size();
        continue;
      }
    }
    if (c < 0x80) {   // ASCII: starters with a fixed class
      flush_run();
      const uint8_t k = t.ascii_cls[c];
      if (k == 3) {
        s.word.push_back(t.ascii_out[c]);
        ++s.word_chars;
      } else if (k == 1) {
        flush_word();
      } else if (k == 2) {
        flush_word();
        if (full()) break;
        const char ch = (char)c;
        wordpiece(t, &ch, 1, 1, s.piece, out);
      }
      ++i;
      continue;
    }
    uint32_t cp;
    const int nb = decode_utf8(p + i, len - i, cp);
    if (nb == 0) return false;
    i += nb;
    const uint32_t e = U.at(cp);
    const uint32_t clean = (e >> 2) & 3u, low = t.lower ? (e >> 4) & 7u : (uint32_t)kLowNone;
    if (clean == kCleanRemove) continue;   // removed before NFD: not a reordering barrier
    if (clean == kCleanSpace) {
      emit(0x20);
      continue;
    }
    if (low == kLowSeq) {   // includes the CJK padding where it applies
      const uint32_t* m = kUniMapIndex[e >> 12];
      for (uint32_t k = 0; k < m[2]; ++k) emit(kUniMapPool[m[1] + k]);
    } else if (low == kLowDrop) {
    } else if (low == kLowDropBarrier) {
      flush_run();
    } else if (low == kLowHangul) {
      const uint32_t h = cp - 0xAC00;
      emit(0x1100 + h / 588);
      emit(0x1161 + (h % 588) / 28);
      if (h % 28) emit(0x11A7 + h % 28);
    } else if (clean == kCleanCjk) {
      emit(0x20);
      emit(cp);
      emit(0x20);
    } else {
      emit(cp);
    }
  }
  // a malformed tail must flag the text even when truncation stopped the loop early
  for (int64_t j = i; j < len;) {
    if (p[j] < 0x80) {
      ++j;
      continue;
    }
    uint32_t cp;
    const int nb = decode_utf8(p + j, len - j, cp);
    if (nb == 0) return false;
    j += nb;
  }
  if (!full()) flush_run();
  if (!full()) flush_word();
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
  t->uni = &uni_table();
  for (int c = 0; c < 128; ++c) {
    const uint32_t e = t->uni->at((uint32_t)c);
    const uint32_t clean = (e >> 2) & 3u;
    t->ascii_out[c] = (char)(t->lower && c >= 'A' && c <= 'Z' ? c + 32 : c);
    if (clean == kCleanRemove) t->ascii_cls[c] = 0;
    else if (clean == kCleanSpace || (e & 3u) == kClsSpace) t->ascii_cls[c] = 1;
    else if ((e & 3u) == kClsPunct) t->ascii_cls[c] = 2;
    else t->ascii_cls[c] = 3;
  }
  *out = t;
  return CSS_OK;
}

int css_tokenizer_destroy(css_tokenizer* t) {
  delete t;
  return CSS_OK;
}

int css_tokenizer_add_special(css_tokenizer* t, const char* literal, int32_t id) {
  CSS_REQUIRE(t != nullptr, "tokenizer is NULL");
  CSS_REQUIRE(literal != nullptr && literal[0] != 0, "empty literal");
  CSS_REQUIRE(id >= 0, "negative id");
  t->special.emplace_back(literal, id);
  std::stable_sort(t->special.begin(), t->special.end(),
                   [](const auto& a, const auto& b) { return a.first.size() > b.first.size(); });
  t->special_first[(unsigned char)literal[0]] = true;
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
    Scratch scratch;
    std::vector<int32_t>& ids = part[w];
    std::vector<int32_t>& ln = part_len[w];
    ids.reserve((size_t)(i1 - i0) * 64);
    ln.reserve(i1 - i0);
    for (int32_t i = i0; i < i1; ++i) {
      const size_t before = ids.size();
      const bool ok = encode_one(*t, texts[i], lens[i], max_len, scratch, ids);
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
