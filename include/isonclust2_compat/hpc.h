// hpc.h of the reference's header face (reference src/hpc.h:6-7, hpc.cpp:4-62).
#ifndef HPC_H_INCLUDED
#define HPC_H_INCLUDED

#include "icb_compat.h"
#include "seq.h"

// homopolymer compression on the device; the compressed read keeps the name, the score and the error rate are left at 0
inline Seq HomopolymerCompressObj(const Seq& s)
{
    const std::string& seq = s.Str();
    const std::string& qual = s.Qual();
    std::string os(seq.size(), '\0'), oq(seq.size(), '\0');
    int32_t len = 0;
    if (!seq.empty()) {
	const int64_t offs[2] = {0, int64_t(seq.size())};
	icb_compat::check(icb_hpc_compress(icb_compat::ctx(), seq.data(), qual.data(), offs, 1, &os[0], &oq[0], &len), "HomopolymerCompress");
    }
    os.resize(size_t(len));
    oq.resize(size_t(len));
    return Seq(s.Name(), os, oq, 0.0);
}

// caller-owned, like the reference's
inline Seq* HomopolymerCompress(const std::unique_ptr<Seq>& s) { return new Seq(HomopolymerCompressObj(*s)); }

#endif
