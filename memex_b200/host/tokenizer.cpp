// tokenizer.cpp -- BERT WordPiece tokenizer, windowing and the segmenter, in C++.
// Mirrors what reference lib/libmemex/src/llm/embedding.rs:155-198 gets from the `tokenizers` crate (0.14.0) for the
// sentence-transformers MiniLM models: BertNormalizer -> BertPreTokenizer -> WordPiece, WordPiece decoder with
// cleanup, truncation with stride.  Checked against the `tokenizers` Python package in tests/test_host_cpp.py.
#include <algorithm>
#include <fstream>

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

std::vector<std::vector<int32_t>> BertTokenizer::encode_windows(const std::string &text, size_t max_length, size_t stride) const
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

std::vector<std::string> segment_text(const ModelConfig &model_config, const std::string &text, const BertTokenizer &tokenizer)
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

TokenBatch tokenize_batch(const BertTokenizer &tokenizer, const std::vector<std::string> &segments, size_t max_seq_length)
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
