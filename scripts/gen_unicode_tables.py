"""Generate claude_semantic_search_b200/csrc/unicode_tables.inc for the native WordPiece tokenizer.

The reference tokenises through sentence-transformers -> AutoTokenizer -> MPNetTokenizerFast, i.e. the
Rust `tokenizers` pipeline BertNormalizer(clean_text, handle_chinese_chars, strip_accents=lowercase,
lowercase) -> BertPreTokenizer -> WordPiece.  Every normaliser step is a per-code-point map except
NFD's canonical reordering, so the tables are taken EMPIRICALLY from that pipeline, one code point at
a time, with the `tokenizers` package of this image (version recorded in the output header):

  * clean-text removals, whitespace -> ' ', CJK padding            (lowercase=False normaliser)
  * NFD + Mn stripping + per-char lower-casing                      (lowercase=True normaliser)
      - identity / dropped (with or without being a reordering barrier) / Hangul (algorithmic,
        verified here for all 11172 syllables) / explicit sequence
  * pre-tokeniser class of every code point: word / whitespace / punctuation
  * canonical combining class of the few non-Mn combining characters (they survive Mn stripping,
    so their NFD reordering is visible in the output)

Run:  python scripts/gen_unicode_tables.py   (rewrites the .inc; tests/test_host_cpu.py sweeps every
code point through both implementations, so a stale table fails the CPU suite).
"""
import sys
import unicodedata as ud
from pathlib import Path

import tokenizers
from tokenizers.normalizers import BertNormalizer
from tokenizers.pre_tokenizers import BertPreTokenizer

OUT = Path(__file__).resolve().parent.parent / "claude_semantic_search_b200" / "csrc" / "unicode_tables.inc"


def ranges(cps):
    out, lo, prev = [], None, None
    for c in sorted(cps):
        if lo is None:
            lo = prev = c
        elif c == prev + 1:
            prev = c
        else:
            out.append((lo, prev))
            lo = prev = c
    if lo is not None:
        out.append((lo, prev))
    return out


def emit_ranges(fh, name, rs):
    fh.write(f"static const uint32_t {name}[][2] = {{\n")
    line = " "
    for lo, hi in rs:
        item = f" {{0x{lo:X},0x{hi:X}}},"
        if len(line) + len(item) > 110:
            fh.write(line + "\n")
            line = " "
        line += item
    fh.write(line + "\n};\n")
    fh.write(f"static const size_t {name}Count = {len(rs)};\n\n")


def hangul_nfd(cp):
    s = cp - 0xAC00
    seq = [0x1100 + s // 588, 0x1161 + (s % 588) // 28]
    if s % 28:
        seq.append(0x11A7 + s % 28)
    return seq


def main():
    plain = BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=False)
    lower = BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=True)
    pre = BertPreTokenizer()
    clean_rm, to_space, cjk = set(), set(), set()
    mn_drop, mn_barrier, hangul, mapped = set(), set(), set(), {}
    pre_space, pre_punct = set(), set()
    # two non-Mn combining characters whose NFD order swaps unless a starter stands between them
    hi_ccc, lo_ccc = "〮", "\U0001d165"
    assert [ord(c) for c in lower.normalize_str(hi_ccc + lo_ccc)] == [0x1D165, 0x302E]
    for cp in range(0x110000):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        c = chr(cp)
        p = plain.normalize_str(c)
        if p == "":
            clean_rm.add(cp)
        elif p == " " and c != " ":
            to_space.add(cp)
        elif p == " " + c + " ":
            cjk.add(cp)
        else:
            assert p == c, (hex(cp), p)
        lw = lower.normalize_str(c)
        if p == "":
            assert lw == ""
        elif p == " " and c != " ":
            assert lw == " "
        elif lw == p:
            pass
        elif lw == "":
            swapped = lower.normalize_str(hi_ccc + c + lo_ccc)
            assert len(swapped) == 2
            (mn_drop if ord(swapped[0]) == 0x1D165 else mn_barrier).add(cp)
        elif 0xAC00 <= cp <= 0xD7A3 and [ord(x) for x in lw] == hangul_nfd(cp):
            hangul.add(cp)
        else:
            mapped[cp] = [ord(x) for x in lw]
        toks = [t for t, _ in pre.pre_tokenize_str("a" + c + "b")]
        if toks == ["a", "b"]:
            pre_space.add(cp)
        elif toks == ["a", c, "b"]:
            pre_punct.add(cp)
        else:
            assert toks == ["a" + c + "b"], (hex(cp), toks)
    assert hangul == set(range(0xAC00, 0xD7A4)), len(hangul)
    def nonstarter(cp):   # survives Mn stripping and does not block the hi/lo swap around it
        got = [ord(x) for x in lower.normalize_str(hi_ccc + chr(cp) + lo_ccc)]
        return len(got) == 3 and got.index(0x1D165) < got.index(0x302E)

    ccc = [(cp, ud.combining(chr(cp))) for cp in range(0x110000)
           if ud.combining(chr(cp)) and ud.category(chr(cp)) != "Mn" and
           (cp in (0x302E, 0x1D165) or nonstarter(cp))]
    # every surviving non-starter is in that list: check the reordering the list implies on the real pipeline
    for a, ca in ccc:
        for b, cb in ccc:
            got = [ord(x) for x in lower.normalize_str(chr(a) + chr(b))]
            want = [b, a] if cb < ca else [a, b]
            assert got == want, (hex(a), hex(b), got)
    with open(OUT, "w", encoding="ascii") as fh:
        fh.write("// GENERATED by scripts/gen_unicode_tables.py -- do not edit.\n")
        fh.write(f"// Source of truth: tokenizers {tokenizers.__version__} (BertNormalizer / BertPreTokenizer), probed one code point\n")
        fh.write(f"// at a time; combining classes from Python unicodedata {ud.unidata_version}.\n\n")
        emit_ranges(fh, "kUniCleanRemove", ranges(clean_rm))
        emit_ranges(fh, "kUniToSpace", ranges(to_space))
        emit_ranges(fh, "kUniCjk", ranges(cjk))
        emit_ranges(fh, "kUniMarkDrop", ranges(mn_drop))
        emit_ranges(fh, "kUniMarkDropBarrier", ranges(mn_barrier))
        emit_ranges(fh, "kUniPreSpace", ranges(pre_space))
        emit_ranges(fh, "kUniPrePunct", ranges(pre_punct))
        emit_ranges(fh, "kUniCcc", ccc)
        pool, index = [], []
        for cp in sorted(mapped):
            index.append((cp, len(pool), len(mapped[cp])))
            pool.extend(mapped[cp])
        fh.write("static const uint32_t kUniMapIndex[][3] = {\n")
        line = " "
        for cp, off, ln in index:
            item = f" {{0x{cp:X},{off},{ln}}},"
            if len(line) + len(item) > 110:
                fh.write(line + "\n")
                line = " "
            line += item
        fh.write(line + "\n};\n")
        fh.write(f"static const size_t kUniMapIndexCount = {len(index)};\n\n")
        fh.write("static const uint32_t kUniMapPool[] = {\n")
        line = " "
        for v in pool:
            item = f" 0x{v:X},"
            if len(line) + len(item) > 110:
                fh.write(line + "\n")
                line = " "
            line += item
        fh.write(line + "\n};\n")
    print(f"{OUT}: clean_rm {len(clean_rm)}, to_space {len(to_space)}, cjk {len(cjk)}, mark_drop {len(mn_drop)}, "
          f"mark_barrier {len(mn_barrier)}, mapped {len(mapped)}, pre_space {len(pre_space)}, "
          f"pre_punct {len(pre_punct)}, ccc {len(ccc)}", file=sys.stderr)


if __name__ == "__main__":
    main()
