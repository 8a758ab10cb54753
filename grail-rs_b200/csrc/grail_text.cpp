// Host-side text front-end, batched and multi-threaded: the reference's Transcriber
// (src/lib.rs:1098-1207) restated over UTF-8 buffers, one text per task.  Runs on the host by design (per character,
// branchy, tiny); its output (phoneme ids + utterance offsets) is exactly what grail_cuda_plan_create_phonemes takes,
// so text-in workloads never materialise per-phoneme records on the host.  No CUDA in this file.
#include "../../include/grail_cuda.h"

#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// UTF-8 -> Unicode scalar values (Rust `str::chars`).  Rust strings are valid UTF-8 by construction; C callers can
// hand us anything, so malformed bytes decode to U+FFFD one byte at a time (what String::from_utf8_lossy would do
// for an isolated bad byte) instead of being trusted.
static void decode_utf8(const char* s, size_t n, std::u32string& out)
{
    out.clear();
    out.reserve(n);
    const unsigned char* p = reinterpret_cast<const unsigned char*>(s);
    size_t i = 0;
    while (i < n) {
        const unsigned c = p[i];
        unsigned need = 0;
        char32_t cp = 0;
        if (c < 0x80) { out.push_back(c); ++i; continue; }
        else if (c >= 0xC2 && c <= 0xDF) { need = 1; cp = c & 0x1F; }
        else if (c >= 0xE0 && c <= 0xEF) { need = 2; cp = c & 0x0F; }
        else if (c >= 0xF0 && c <= 0xF4) { need = 3; cp = c & 0x07; }
        else { out.push_back(0xFFFD); ++i; continue; }
        bool ok = i + need < n;                          // all continuation bytes are inside the buffer
        if (ok)
            for (unsigned k = 1; k <= need; ++k) {
                if ((p[i + k] & 0xC0) != 0x80) { ok = false; break; }
                cp = (cp << 6) | (p[i + k] & 0x3F);
            }
        if (ok && ((need == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) || (need == 3 && (cp < 0x10000 || cp > 0x10FFFF))))
            ok = false;
        if (!ok) { out.push_back(0xFFFD); ++i; continue; }
        out.push_back(cp);
        i += need + 1;
    }
}

struct Rule {
    std::u32string string;
    const uint8_t* phonemes;
    uint32_t n_phonemes;
};

// `x.string.chars().nth(index)` (:1145, :1151): the index-th scalar or None
static inline bool nth(const Rule& r, size_t index, char32_t& c)
{
    if (index >= r.string.size()) return false;
    c = r.string[index];
    return true;
}

template <class Pred>
static size_t partition_point(const std::vector<Rule>& rules, size_t lo, size_t hi, Pred pred)
{
    while (lo < hi) {                                   // slice::partition_point: first index where pred is false
        const size_t mid = lo + (hi - lo) / 2;
        if (pred(rules[mid])) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static const uint8_t SILENCE_BUF[1] = { GRAIL_PHONEME_SILENCE };                        // :1115

// Drains one Transcriber (Iterator::next until None, :1117-1189) over `text`.
template <class Emit>
static void transcribe_one(const std::u32string& text, const std::vector<Rule>& rules, bool case_sensitive,
                           bool leading_silence, Emit emit)
{
    size_t pos = 0;                                     // Peekable<Chars>: text[pos] is what peek() sees
    const uint8_t* buf = leading_silence ? SILENCE_BUF : nullptr;                       // :1201 (tests start empty)
    size_t buf_n = leading_silence ? 1 : 0;
    for (;;) {
        size_t search_min = 0, search_max = rules.size(), index = 0;                    // :1121-1123
        while (buf_n == 0) {                                                            // :1126
            if (pos >= text.size()) return;                                             // peek() is None: `?` (:1134)
            char32_t ch = text[pos];
            if (!case_sensitive && ch >= U'A' && ch <= U'Z') ch += 32;                  // to_ascii_lowercase (:1131)
            const size_t new_min = partition_point(rules, search_min, search_max, [&](const Rule& r) {
                char32_t c;
                return !nth(r, index, c) || c < ch;                                     // map_or(true, |x| x < character)
            });
            const size_t new_max = partition_point(rules, search_min, search_max, [&](const Rule& r) {
                char32_t c;
                return nth(r, index, c) && c <= ch;                                     // map_or(false, |x| x <= character)
            });
            // `self.ruleset[search_min]` panics in the reference when the rule set is empty; here: no rule matches
            const bool min_matched = search_min < rules.size() && rules[search_min].string.size() == index;
            if (new_min >= new_max && min_matched) {                                    // :1158-1160
                buf = rules[search_min].phonemes;
                buf_n = rules[search_min].n_phonemes;
            } else if (new_min >= new_max) {                                            // :1161-1166
                buf = SILENCE_BUF;
                buf_n = 1;
                ++pos;
            } else {                                                                    // :1167-1184
                search_min = new_min;
                search_max = new_max;
                ++index;
                ++pos;
                if (pos >= text.size()) {
                    if (rules[search_min].string.size() == index) {
                        buf = rules[search_min].phonemes;
                        buf_n = rules[search_min].n_phonemes;
                    } else {
                        buf = SILENCE_BUF;
                        buf_n = 1;
                    }
                }
            }
        }
        emit(buf[0]);                                                                   // :1187-1192
        ++buf;
        --buf_n;
    }
}
// Two rule shapes make the reference loop forever without consuming input and are rejected by
// grail_cuda_transcribe_batch instead: an empty phoneme list (the buffer stays empty at :1160 with the range unable
// to narrow) and an empty rule string (it "matches" at index 0, :1158, before any character is taken).

}  // namespace

extern "C" int grail_cuda_transcribe_batch(const char* const* texts, const size_t* text_bytes, uint32_t n_texts,
                                           const grail_transcription_rule* rules, uint32_t n_rules, int case_sensitive,
                                           int leading_silence, uint8_t* ids, uint64_t ids_capacity,
                                           uint32_t* utt_offsets, int n_threads)
{
    if ((n_texts && !texts) || (n_rules && !rules) || !utt_offsets) return GRAIL_ERR_INVALID_ARG;
    std::vector<Rule> rs(n_rules);
    for (uint32_t i = 0; i < n_rules; ++i) {
        if (!rules[i].string || !rules[i].string[0] || !rules[i].phonemes || rules[i].n_phonemes == 0) return GRAIL_ERR_INVALID_ARG;
        decode_utf8(rules[i].string, strlen(rules[i].string), rs[i].string);
        rs[i].phonemes = rules[i].phonemes;
        rs[i].n_phonemes = rules[i].n_phonemes;
        if (i && !(rs[i - 1].string <= rs[i].string)) return GRAIL_ERR_INVALID_ARG;     // "assumed to be sorted" (:1095)
    }
    for (uint32_t t = 0; t < n_texts; ++t)
        if (!texts[t]) return GRAIL_ERR_INVALID_ARG;
    unsigned T = n_threads > 0 ? (unsigned)n_threads : std::thread::hardware_concurrency();
    if (T == 0) T = 1;
    if (T > n_texts) T = n_texts ? n_texts : 1;

    std::vector<std::vector<uint8_t>> out(n_texts);
    std::atomic<uint32_t> next(0);
    auto worker = [&]() {
        std::u32string text;
        for (;;) {
            const uint32_t t0 = next.fetch_add(16);                                     // 16 texts per grab
            if (t0 >= n_texts) return;
            const uint32_t t1 = t0 + 16 < n_texts ? t0 + 16 : n_texts;
            for (uint32_t t = t0; t < t1; ++t) {
                decode_utf8(texts[t], text_bytes ? text_bytes[t] : strlen(texts[t]), text);
                std::vector<uint8_t>& o = out[t];
                o.reserve(text.size() + 1);
                transcribe_one(text, rs, case_sensitive != 0, leading_silence != 0, [&o](uint8_t p) { o.push_back(p); });
            }
        }
    };
    if (T == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        pool.reserve(T);
        for (unsigned i = 0; i < T; ++i) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    uint64_t total = 0;
    utt_offsets[0] = 0;
    for (uint32_t t = 0; t < n_texts; ++t) {
        total += out[t].size();
        if (total > 0xFFFFFFFFull) return GRAIL_ERR_UNSUPPORTED;
        utt_offsets[t + 1] = (uint32_t)total;
    }
    if (!ids) return GRAIL_OK;                                                          // counting call
    if (ids_capacity < total) return GRAIL_ERR_COUNT_MISMATCH;
    for (uint32_t t = 0; t < n_texts; ++t)
        if (!out[t].empty()) memcpy(ids + utt_offsets[t], out[t].data(), out[t].size());
    return GRAIL_OK;
}
