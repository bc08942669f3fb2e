// tokenizer.cpp -- BERT WordPiece and byte-level BPE tokenizers, windowing and the segmenter, in C++.
// Mirrors what reference lib/libmemex/src/llm/embedding.rs:155-198 gets from the `tokenizers` crate (0.14.0) for the
// three models it segments: the MiniLM models (BertNormalizer -> BertPreTokenizer -> WordPiece, WordPiece decoder with
// cleanup) and all-distilroberta-v1 (ByteLevel pre-tokenizer -> BPE -> ByteLevel decoder, RobertaProcessing), plus
// truncation with stride.  Checked against the `tokenizers` Python package in tests/test_host_cpp.py.
#include <algorithm>
#include <fstream>

#include <climits>
#include <iterator>

#include "json.hpp"
#include "memex_host.hpp"
#include "unicode_tables.hpp"

namespace memex {

namespace {

template <size_t N>
bool in_ranges(const uint16_t (&tab)[N][2], char32_t c)
{
    if (c > 0xFFFF) return false;
    size_t lo = 0, hi = N;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (c < tab[mid][0]) hi = mid;
        else if (c > tab[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}

std::u32string from_utf8(const std::string &s)
{
    std::u32string out;
    out.reserve(s.size());
    size_t i = 0;
    while (i < s.size()) {
        unsigned char c = (unsigned char)s[i];
        char32_t cp;
        size_t n;
        if (c < 0x80) { cp = c; n = 1; }
        else if ((c >> 5) == 6) { cp = c & 0x1F; n = 2; }
        else if ((c >> 4) == 14) { cp = c & 0x0F; n = 3; }
        else if ((c >> 3) == 30) { cp = c & 0x07; n = 4; }
        else { cp = 0xFFFD; n = 1; }
        if (i + n > s.size()) { out.push_back(0xFFFD); break; }
        for (size_t j = 1; j < n; ++j) {
            unsigned char cc = (unsigned char)s[i + j];
            if ((cc >> 6) != 2) { cp = 0xFFFD; n = j; break; }
            cp = (cp << 6) | (cc & 0x3F);
        }
        out.push_back(cp);
        i += n;
    }
    return out;
}

void append_utf8(char32_t cp, std::string &out)
{
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
}

std::string to_utf8(const std::u32string &s)
{
    std::string out;
    for (char32_t c : s) append_utf8(c, out);
    return out;
}

bool is_whitespace(char32_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || in_ranges(utab::kSpace, c); }
bool is_control(char32_t c)
{
    if (c == '\t' || c == '\n' || c == '\r') return false;
    return in_ranges(utab::kOther, c);
}
bool is_punctuation(char32_t c)
{
    // BERT treats all non-letter / non-digit ASCII as punctuation, whatever its Unicode class ("^", "$", "`" ...)
    if ((c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126)) return true;
    return in_ranges(utab::kPunct, c);
}
bool is_chinese_char(char32_t c)
{
    return (c >= 0x4E00 && c <= 0x9FFF) || (c >= 0x3400 && c <= 0x4DBF) || (c >= 0x20000 && c <= 0x2A6DF) ||
           (c >= 0x2A700 && c <= 0x2B73F) || (c >= 0x2B740 && c <= 0x2B81F) || (c >= 0x2B920 && c <= 0x2CEAF) ||
           (c >= 0xF900 && c <= 0xFAFF) || (c >= 0x2F800 && c <= 0x2FA1F);
}

// NFD with the non-spacing marks dropped
void strip_accents_into(char32_t c, std::u32string &out)
{
    if (c < 0x80) { out.push_back(c); return; }
    if (in_ranges(utab::kMn, c)) return;
    if (c >= 0xAC00 && c <= 0xD7A3) {   // Hangul syllable -> jamo
        const uint32_t s = c - 0xAC00;
        out.push_back(0x1100 + s / 588);
        out.push_back(0x1161 + (s % 588) / 28);
        if (s % 28) out.push_back(0x11A7 + s % 28);
        return;
    }
    if (c <= 0xFFFF) {
        constexpr size_t n = sizeof(utab::kDecomp) / sizeof(utab::kDecomp[0]);
        size_t lo = 0, hi = n;
        while (lo < hi) {
            size_t mid = (lo + hi) / 2;
            if (utab::kDecomp[mid].cp < c) lo = mid + 1;
            else hi = mid;
        }
        if (lo < n && utab::kDecomp[lo].cp == c) {
            for (int i = 0; i < utab::kDecomp[lo].n; ++i) out.push_back(utab::kDecomp[lo].c[i]);
            return;
        }
    }
    out.push_back(c);
}

char32_t to_lower(char32_t c)
{
    if (c < 0x80) return (c >= 'A' && c <= 'Z') ? c + 32 : c;
    if (c > 0xFFFF) return c;
    constexpr size_t n = sizeof(utab::kLower) / sizeof(utab::kLower[0]);
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (utab::kLower[mid][0] < c) lo = mid + 1;
        else hi = mid;
    }
    return (lo < n && utab::kLower[lo][0] == c) ? (char32_t)utab::kLower[lo][1] : c;
}

// decoders::wordpiece::cleanup of the `tokenizers` crate, applied per token
std::string cleanup(std::string s)
{
    static const char *const rep[][2] = {{" .", "."}, {" ?", "?"}, {" !", "!"}, {" ,", ","}, {" ' ", "'"}, {" n't", "n't"},
                                         {" 'm", "'m"}, {" do not", " don't"}, {" 's", "'s"}, {" 've", "'ve"}, {" 're", "'re"}};
    for (const auto &r : rep) {
        const std::string from = r[0], to = r[1];
        size_t pos = 0;
        while ((pos = s.find(from, pos)) != std::string::npos) {
            s.replace(pos, from.size(), to);
            pos += to.size();
        }
    }
    return s;
}

}  // namespace

std::shared_ptr<BertTokenizer> BertTokenizer::from_vocab(const std::vector<std::string> &tokens, bool lowercase)
{
    auto t = std::make_shared<BertTokenizer>();
    t->lowercase_ = lowercase;
    t->id_to_token_ = tokens;
    for (size_t i = 0; i < tokens.size(); ++i) t->token_to_id_.emplace(tokens[i], (int32_t)i);
    auto id_of = [&](const char *name, int32_t fallback) {
        auto it = t->token_to_id_.find(name);
        return it == t->token_to_id_.end() ? fallback : it->second;
    };
    t->pad_id = id_of("[PAD]", 0);
    t->unk_id = id_of("[UNK]", 100);
    t->cls_id = id_of("[CLS]", 101);
    t->sep_id = id_of("[SEP]", 102);
    t->mask_id = id_of("[MASK]", 103);
    return t;
}

std::shared_ptr<BertTokenizer> BertTokenizer::from_vocab_file(const std::string &vocab_txt, bool lowercase)
{
    std::ifstream f(vocab_txt);
    if (!f) throw EmbeddingError(EmbeddingErrorKind::SetupError, "Unable to load model <" + vocab_txt + ">");
    std::vector<std::string> tokens;
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        tokens.push_back(line);
    }
    return from_vocab(tokens, lowercase);
}

// BertNormalizer (clean_text, handle_chinese_chars, strip_accents = lowercase, lowercase) + BertPreTokenizer
std::vector<std::u32string> BertTokenizer::pre_tokenize(const std::string &text) const
{
    const std::u32string raw = from_utf8(text);
    std::u32string norm;
    norm.reserve(raw.size() + 16);
    for (char32_t c : raw) {
        if (c == 0 || c == 0xFFFD || is_control(c)) continue;          // clean_text
        if (is_whitespace(c)) { norm.push_back(' '); continue; }
        if (is_chinese_char(c)) {                                      // handle_chinese_chars
            norm.push_back(' ');
            norm.push_back(c);
            norm.push_back(' ');
            continue;
        }
        if (lowercase_) {
            std::u32string d;
            strip_accents_into(c, d);
            for (char32_t x : d) norm.push_back(to_lower(x));
        } else {
            norm.push_back(c);
        }
    }
    std::vector<std::u32string> words;
    std::u32string cur;
    for (char32_t c : norm) {
        if (c == ' ' || is_whitespace(c)) {
            if (!cur.empty()) { words.push_back(cur); cur.clear(); }
        } else if (is_punctuation(c)) {
            if (!cur.empty()) { words.push_back(cur); cur.clear(); }
            words.push_back(std::u32string(1, c));
        } else {
            cur.push_back(c);
        }
    }
    if (!cur.empty()) words.push_back(cur);
    return words;
}

// WordPiece: greedy longest match first, "##" continuation prefix, max 100 characters per word
void BertTokenizer::wordpiece(const std::u32string &word, std::vector<int32_t> &out) const
{
    if (word.size() > 100) { out.push_back(unk_id); return; }
    std::vector<int32_t> pieces;
    size_t start = 0;
    while (start < word.size()) {
        size_t end = word.size();
        int32_t found = -1;
        while (end > start) {
            std::string piece = start > 0 ? "##" : "";
            piece += to_utf8(word.substr(start, end - start));
            auto it = token_to_id_.find(piece);
            if (it != token_to_id_.end()) { found = it->second; break; }
            --end;
        }
        if (found < 0) { out.push_back(unk_id); return; }
        pieces.push_back(found);
        start = end;
    }
    out.insert(out.end(), pieces.begin(), pieces.end());
}

std::vector<int32_t> BertTokenizer::encode(const std::string &text, bool add_special_tokens) const
{
    std::vector<int32_t> ids;
    if (add_special_tokens) ids.push_back(cls_id);
    for (const auto &w : pre_tokenize(text)) wordpiece(w, ids);
    if (add_special_tokens) ids.push_back(sep_id);
    return ids;
}

std::string BertTokenizer::decode(const std::vector<int32_t> &ids, bool skip_special_tokens) const
{
    std::string out;
    bool first = true;
    for (int32_t id : ids) {
        if (id < 0 || (size_t)id >= id_to_token_.size()) continue;
        if (skip_special_tokens && (id == pad_id || id == unk_id || id == cls_id || id == sep_id || id == mask_id)) continue;
        std::string tok = id_to_token_[id];
        if (!first) {
            if (tok.rfind("##", 0) == 0) tok = tok.substr(2);
            else tok = " " + tok;
        }
        out += cleanup(tok);
        first = false;
    }
    return out;
}

std::vector<std::vector<int32_t>> Tokenizer::encode_windows(const std::string &text, size_t max_length, size_t stride) const
{
    const std::vector<int32_t> ids = encode(text, false);
    std::vector<std::vector<int32_t>> out;
    if (max_length == 0 || stride >= max_length) throw EmbeddingError(EmbeddingErrorKind::SetupError, "stride must be smaller than max_length");
    if (ids.size() <= max_length) {
        out.push_back(ids);
        return out;
    }
    const size_t step = max_length - stride;
    for (size_t start = 0;; start += step) {
        const size_t end = std::min(ids.size(), start + max_length);
        out.emplace_back(ids.begin() + start, ids.begin() + end);
        if (end == ids.size()) break;
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// byte-level BPE (RoBERTa)
// ------------------------------------------------------------------------------------------------
namespace {

template <size_t N>
bool in_ranges32(const uint32_t (&tab)[N][2], char32_t c)
{
    size_t lo = 0, hi = N;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (c < tab[mid][0]) hi = mid;
        else if (c > tab[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}
bool is_letter(char32_t c) { return in_ranges32(utab::kLetter32, c); }   // \p{L}
bool is_number(char32_t c) { return in_ranges32(utab::kNumber32, c); }   // \p{N}
// \s of the pre-tokenizer's regex engine: the Unicode White_Space property (pinned by the golden vectors: U+001C-1F,
// U+200B, U+180E and U+FEFF are NOT whitespace there)
bool is_regex_space(char32_t c)
{
    return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) ||
           c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}

// GPT-2's bytes_to_unicode: printable bytes map to themselves, the others to U+0100 + n in order
struct ByteMap {
    char32_t to_cp[256];
    std::unordered_map<char32_t, uint8_t> to_byte;
    ByteMap()
    {
        int n = 0;
        for (int b = 0; b < 256; ++b) {
            const bool keep = (b >= 0x21 && b <= 0x7E) || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE && b <= 0xFF);
            to_cp[b] = keep ? (char32_t)b : (char32_t)(0x100 + n++);
            to_byte.emplace(to_cp[b], (uint8_t)b);
        }
    }
};
const ByteMap &byte_map()
{
    static const ByteMap m;
    return m;
}

// 's|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+   -> [begin, end) pieces
void gpt2_split(const std::u32string &t, std::vector<std::pair<size_t, size_t>> &out)
{
    const size_t n = t.size();
    size_t i = 0;
    auto other = [](char32_t c) { return !is_regex_space(c) && !is_letter(c) && !is_number(c); };
    while (i < n) {
        size_t end = 0;
        if (t[i] == U'\'' && i + 1 < n) {
            const char32_t a = t[i + 1], b = i + 2 < n ? t[i + 2] : 0;
            if (a == U's' || a == U't' || a == U'm' || a == U'd') end = i + 2;
            else if ((a == U'r' && b == U'e') || (a == U'v' && b == U'e') || (a == U'l' && b == U'l')) end = i + 3;
        }
        if (!end) {
            const size_t j = (t[i] == U' ' && i + 1 < n) ? i + 1 : i;   // the optional single space
            size_t k = j;
            if (is_letter(t[j])) { while (k < n && is_letter(t[k])) ++k; }
            else if (is_number(t[j])) { while (k < n && is_number(t[k])) ++k; }
            else if (other(t[j])) { while (k < n && other(t[k])) ++k; }
            if (k > j) end = k;
        }
        if (!end) {   // whitespace run; leave its last character to the next piece when a non-space follows
            size_t k = i;
            while (k < n && is_regex_space(t[k])) ++k;
            end = (k < n && k - i > 1) ? k - 1 : k;
        }
        out.emplace_back(i, end);
        i = end;
    }
}

// String::from_utf8_lossy: every maximal invalid subsequence becomes one U+FFFD
std::string utf8_lossy(const std::string &b)
{
    std::string out;
    size_t i = 0;
    const size_t n = b.size();
    auto cont = [&](size_t p, unsigned lo = 0x80, unsigned hi = 0xBF) { return p < n && (unsigned char)b[p] >= lo && (unsigned char)b[p] <= hi; };
    while (i < n) {
        const unsigned char c = (unsigned char)b[i];
        size_t len = 0, bad = 1;
        if (c < 0x80) len = 1;
        else if (c >= 0xC2 && c <= 0xDF) { if (cont(i + 1)) len = 2; }
        else if (c >= 0xE0 && c <= 0xEF) {
            const unsigned lo = c == 0xE0 ? 0xA0 : 0x80, hi = c == 0xED ? 0x9F : 0xBF;
            if (cont(i + 1, lo, hi)) { if (cont(i + 2)) len = 3; else bad = 2; }
        } else if (c >= 0xF0 && c <= 0xF4) {
            const unsigned lo = c == 0xF0 ? 0x90 : 0x80, hi = c == 0xF4 ? 0x8F : 0xBF;
            if (cont(i + 1, lo, hi)) {
                if (cont(i + 2)) { if (cont(i + 3)) len = 4; else bad = 3; }
                else bad = 2;
            }
        }
        if (len) { out.append(b, i, len); i += len; }
        else { out += "\xEF\xBF\xBD"; i += bad; }
    }
    return out;
}

}  // namespace

std::shared_ptr<ByteLevelBpeTokenizer> ByteLevelBpeTokenizer::from_vocab(const std::vector<std::string> &tokens,
                                                                         const std::vector<std::pair<std::string, std::string>> &merges)
{
    auto t = std::make_shared<ByteLevelBpeTokenizer>();
    t->id_to_token_ = tokens;
    for (size_t i = 0; i < tokens.size(); ++i) t->token_to_id_.emplace(tokens[i], (int32_t)i);
    for (size_t r = 0; r < merges.size(); ++r) t->merge_rank_.emplace(merges[r].first + " " + merges[r].second, (uint32_t)r);
    auto id_of = [&](const char *name, int32_t fallback) {
        auto it = t->token_to_id_.find(name);
        return it == t->token_to_id_.end() ? fallback : it->second;
    };
    t->cls_id = id_of("<s>", 0);
    t->pad_id = id_of("<pad>", 1);
    t->sep_id = id_of("</s>", 2);
    t->unk_id = id_of("<unk>", 3);
    t->mask_id = id_of("<mask>", -1);
    return t;
}

std::shared_ptr<ByteLevelBpeTokenizer> ByteLevelBpeTokenizer::from_files(const std::string &vocab_json, const std::string &merges_txt)
{
    std::ifstream vf(vocab_json, std::ios::binary), mf(merges_txt, std::ios::binary);
    if (!vf) throw EmbeddingError(EmbeddingErrorKind::SetupError, "Unable to load model <" + vocab_json + ">");
    if (!mf) throw EmbeddingError(EmbeddingErrorKind::SetupError, "Unable to load model <" + merges_txt + ">");
    std::string text((std::istreambuf_iterator<char>(vf)), std::istreambuf_iterator<char>());
    json::Value v;
    try {
        v = json::parse(text);
    } catch (const std::exception &e) {
        throw EmbeddingError(EmbeddingErrorKind::SetupError, vocab_json + ": " + e.what());
    }
    std::vector<std::string> tokens(v.obj.size());
    for (const auto &kv : v.obj) {
        const size_t id = (size_t)kv.second.num;
        if (id >= tokens.size()) tokens.resize(id + 1);
        tokens[id] = kv.first;
    }
    std::vector<std::pair<std::string, std::string>> merges;
    std::string line;
    while (std::getline(mf, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line.rfind("#version", 0) == 0) continue;
        const size_t sp = line.find(' ');
        if (sp == std::string::npos) continue;
        merges.emplace_back(line.substr(0, sp), line.substr(sp + 1));
    }
    return from_vocab(tokens, merges);
}

// one pre-token (byte-level spelling) -> ids: merge the adjacent pair of lowest rank (leftmost on ties) until none is left
void ByteLevelBpeTokenizer::bpe(const std::string &piece, std::vector<int32_t> &out) const
{
    std::vector<std::string> sym;
    for (char32_t c : from_utf8(piece)) {
        std::string s;
        append_utf8(c, s);
        sym.push_back(std::move(s));
    }
    while (sym.size() > 1) {
        uint32_t best = UINT32_MAX;
        size_t at = 0;
        for (size_t i = 0; i + 1 < sym.size(); ++i) {
            auto it = merge_rank_.find(sym[i] + " " + sym[i + 1]);
            if (it != merge_rank_.end() && it->second < best) {
                best = it->second;
                at = i;
            }
        }
        if (best == UINT32_MAX) break;
        sym[at] += sym[at + 1];
        sym.erase(sym.begin() + at + 1);
    }
    for (const auto &s : sym) {
        auto it = token_to_id_.find(s);
        // RoBERTa's BPE model has no unk_token: a symbol outside the vocabulary is dropped (cannot happen with the
        // full byte alphabet in the vocabulary)
        if (it != token_to_id_.end()) out.push_back(it->second);
    }
}

std::vector<int32_t> ByteLevelBpeTokenizer::encode(const std::string &text, bool add_special_tokens) const
{
    std::vector<int32_t> ids;
    if (add_special_tokens) ids.push_back(cls_id);
    const std::u32string t = from_utf8(text);
    std::vector<std::pair<size_t, size_t>> pieces;
    gpt2_split(t, pieces);
    const ByteMap &bm = byte_map();
    for (const auto &pc : pieces) {
        const std::string raw = to_utf8(t.substr(pc.first, pc.second - pc.first));
        std::string mapped;
        for (unsigned char b : raw) append_utf8(bm.to_cp[b], mapped);
        bpe(mapped, ids);
    }
    if (add_special_tokens) ids.push_back(sep_id);
    return ids;
}

std::string ByteLevelBpeTokenizer::decode(const std::vector<int32_t> &ids, bool skip_special_tokens) const
{
    const ByteMap &bm = byte_map();
    std::string bytes;
    for (int32_t id : ids) {
        if (id < 0 || (size_t)id >= id_to_token_.size()) continue;
        if (skip_special_tokens && (id == pad_id || id == unk_id || id == cls_id || id == sep_id || id == mask_id)) continue;
        for (char32_t c : from_utf8(id_to_token_[id])) {
            auto it = bm.to_byte.find(c);
            if (it != bm.to_byte.end()) bytes += (char)it->second;
        }
    }
    return utf8_lossy(bytes);
}

std::vector<std::string> segment_text(const ModelConfig &model_config, const std::string &text, const Tokenizer &tokenizer)
{
    switch (model_config.model) {   // embedding.rs:156-161
        case EmbeddingsModelType::AllMiniLmL12V2:
        case EmbeddingsModelType::AllMiniLmL6V2:
        case EmbeddingsModelType::AllDistilrobertaV1:
            break;
        default:
            throw EmbeddingError(EmbeddingErrorKind::SetupError, "Model not supported yet");
    }
    std::vector<std::string> segments;
    const auto windows = tokenizer.encode_windows(text, model_config.max_length, model_config.stride);
    for (size_t i = 0; i < windows.size(); ++i) {
        std::string decoded = tokenizer.decode(windows[i], true);
        if (i == 0) {   // embedding.rs:183: only the first window gets this replacement
            size_t pos = 0;
            while ((pos = decoded.find(" ' ", pos)) != std::string::npos) {
                decoded.replace(pos, 3, "'");
                pos += 1;
            }
        }
        segments.push_back(std::move(decoded));
    }
    return segments;
}

TokenBatch tokenize_batch(const Tokenizer &tokenizer, const std::vector<std::string> &segments, size_t max_seq_length)
{
    TokenBatch tb;
    tb.B = (uint32_t)segments.size();
    std::vector<std::vector<int32_t>> all;
    size_t longest = 1;
    for (const auto &s : segments) {
        std::vector<int32_t> ids = tokenizer.encode(s, false);
        if (max_seq_length >= 2 && ids.size() > max_seq_length - 2) ids.resize(max_seq_length - 2);   // room for [CLS] / [SEP]
        ids.insert(ids.begin(), tokenizer.cls_id);
        ids.push_back(tokenizer.sep_id);
        longest = std::max(longest, ids.size());
        all.push_back(std::move(ids));
    }
    tb.S = (uint32_t)longest;
    tb.ids.assign((size_t)tb.B * tb.S, tokenizer.pad_id);
    tb.lens.resize(tb.B);
    for (size_t i = 0; i < all.size(); ++i) {
        std::copy(all[i].begin(), all[i].end(), tb.ids.begin() + i * tb.S);
        tb.lens[i] = (int32_t)all[i].size();
    }
    return tb;
}

}  // namespace memex
