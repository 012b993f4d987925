// strip.h -- host pre-strip of FASTQ input (strip.cpp).  Internal to libfinch_b200.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace fb2 {

struct StripOut {
    uint8_t *out = nullptr;          // "sequence line\n" per record
    size_t out_cap = 0, out_len = 0;
    const uint8_t *origin = nullptr; // stream offsets are base_off + (p - origin)
    uint64_t bases = 0;              // sum of sequence().len() (CR trimmed)
    uint64_t records = 0;
    uint64_t bad_pos = ~0ull;        // min stream offset of a line start failing the '@' / '+' check (non-blank records)
    uint64_t len_bad_pos = ~0ull;    // min stream offset of the header newline of a record whose seq / qual lengths differ
    uint64_t first_blank = ~0ull;    // min stream offset of a "record" made of line terminators only
    uint64_t last_nonblank = 0;      // max stream offset + 1 of the start of a record with content (0: none)
    std::vector<uint8_t> spill;      // used instead of `out` when a range had to be redone sequentially
    bool spilled = false;
    void reset_counts() {
        out_len = 0; bases = records = 0; bad_pos = len_bad_pos = first_blank = ~0ull; last_nonblank = 0; spilled = false;
    }
    const uint8_t *data() const { return spilled ? spill.data() : out; }
};

// Whole records of [p, end), p at a record start; see strip.cpp.
const uint8_t *strip_records(const uint8_t *p, const uint8_t *end, const uint8_t *stop_at, uint64_t base_off, StripOut &o);
const uint8_t *strip_parallel(const uint8_t *p, const uint8_t *end, uint64_t base_off, unsigned threads,
                              std::vector<StripOut> &outs);
// The end of the stream: `p` .. `end` is what is left after the last complete record (from a record start).
// Applies the reader's end-of-input rules (a last quality line may lack its newline; trailing blank lines are
// fine; anything else is a truncated / invalid record).  Returns 0 = ok, 1 = invalid record.
int strip_final(const uint8_t *p, const uint8_t *end, uint64_t base_off, StripOut &o);
// Offset of the end of the record that starts in `carry` and continues in [p, end): the position right after the
// newline that completes its fourth line, or nullptr when [p, end) does not hold enough newlines.
const uint8_t *complete_record(const std::vector<uint8_t> &carry, const uint8_t *p, const uint8_t *end);

}  // namespace fb2
