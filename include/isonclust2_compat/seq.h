// seq.h of the reference's header face: the read record the path's functions take (reference src/seq.h:21-101).
#ifndef SEQ_H_INCLUDED
#define SEQ_H_INCLUDED

#include <stdint.h>

#include <cstdint>

#include <cmath>
#include <memory>
#include <string>
#include <vector>

class Seq;
typedef std::vector<std::unique_ptr<Seq>> SequencesP;
typedef std::vector<Seq> Sequences;

class Seq {
public:
    Seq() {}
    Seq(const std::string& name, const std::string& seq, const std::string& qual, double score)
	: name_(name), seq_(seq), qual_(qual), score_(score), error_rate_(0.0)
    {
    }
    const std::string& Name() const { return name_; }
    void SetName(const std::string& n) { name_ = n; }
    const std::string& Str() const { return seq_; }
    unsigned Len() const { return unsigned(seq_.length()); }
    void SetStr(const std::string& s) { seq_ = s; }
    const std::string& Qual() const { return qual_; }
    void SetQual(const std::string& q) { qual_ = q; }
    unsigned MeanQual() const { return unsigned(-10 * log10(error_rate_)); }
    std::vector<uint32_t> Weights() const
    {
	std::vector<uint32_t> w(qual_.size());
	for (size_t i = 0; i < qual_.size(); i++) w[i] = uint32_t(qual_[i] - 33);
	return w;
    }
    double Score() const { return score_; }
    void SetScore(double s) { score_ = s; }
    double ErrorRate() const { return error_rate_; }
    void SetErrorRate(double e) { error_rate_ = e; }

private:
    std::string name_, seq_, qual_;
    double score_ = 0.0, error_rate_ = 0.0;
};

#endif
