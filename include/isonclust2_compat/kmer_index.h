// kmer_index.h of the reference's header face (reference src/kmer_index.h:8-66, kmer_index.cpp:5-17).
#ifndef KMER_INDEX_H_INCLUDED
#define KMER_INDEX_H_INCLUDED

#include <string>
#include <vector>

#include "icb_compat.h"

typedef std::vector<unsigned> KmerSeq;

// 2-bit values of the k-mers of seq, on the device (the reference drops the last k-mer: len - k values)
inline KmerSeq KmerEncodeSeq(const std::string seq, unsigned kmerSize)
{
    KmerSeq out(seq.size() > kmerSize ? seq.size() - kmerSize : 0);
    int64_t n = 0;
    icb_compat::check(icb_kmer_encode(icb_compat::ctx(), seq.data(), int64_t(seq.size()), int(kmerSize), out.data(), &n), "KmerEncodeSeq");
    out.resize(size_t(n));
    return out;
}

inline char NumberToBase(unsigned i) { return i < 4 ? "ACGT"[i] : '!'; }

inline unsigned BaseToNumber(char c)
{
    const char* p = "ACGT";
    for (unsigned i = 0; i < 4; i++)
	if (p[i] == c) return i;
    return unsigned(-1);
}

// base-4 digits of index, most significant first
inline std::string IndexToKmer(unsigned index, unsigned k)
{
    std::string s(k, 'A');
    for (unsigned j = k; j-- > 0; index /= 4) s[j] = NumberToBase(index % 4);
    return s;
}

// value of the prefix [begin, end) of kmer read as base-4 digits (unsigned arithmetic, as the reference's recursion)
inline unsigned KmerToIndex(const std::string& kmer, std::string::iterator end)
{
    unsigned v = 0;
    const size_t n = size_t(end - const_cast<std::string&>(kmer).begin());
    for (size_t i = 0; i < n; i++) v = 4 * v + BaseToNumber(kmer[i]);
    return v;
}

#endif
