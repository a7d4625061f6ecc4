// minimap2_b200/csrc/fastx.cc -- see fastx.h
#include "fastx.h"
#include <cctype>
#include <cstring>
#include <unistd.h>

FastxReader::FastxReader(const char *fn)
{
	fp_ = (fn && strcmp(fn, "-") != 0)? gzopen(fn, "r") : gzdopen(0, "r");
	if (fp_) gzbuffer(fp_, 1 << 20);
	buf_.resize(1 << 20);
}

FastxReader::~FastxReader() { if (fp_) gzclose(fp_); }

int FastxReader::getc_()
{
	if (beg_ >= end_) {
		if (eof_) return -1;
		int n = gzread(fp_, buf_.data(), (unsigned)buf_.size());
		if (n <= 0) { eof_ = true; return -1; }
		beg_ = 0, end_ = (size_t)n;
	}
	return buf_[beg_++];
}

int FastxReader::peek_()
{
	int c = getc_();
	if (c >= 0) --beg_;
	return c;
}

bool FastxReader::eof()
{
	if (last_) return false;
	return peek_() < 0;
}

void FastxReader::readline_(std::string &s, bool append)
{
	if (!append) s.clear();
	for (;;) {
		if (beg_ >= end_) { if (getc_() < 0) return; --beg_; }
		unsigned char *p = buf_.data() + beg_, *e = buf_.data() + end_;
		unsigned char *nl = (unsigned char*)memchr(p, '\n', e - p);
		if (nl) {
			s.append((char*)p, nl - p);
			beg_ = (nl - buf_.data()) + 1;
			break;
		}
		s.append((char*)p, e - p);
		beg_ = end_;
	}
	while (!s.empty() && s.back() == '\r') s.pop_back();
}

int FastxReader::next(FastxRecord &r, bool with_qual, bool with_comment)
{
	int c;
	if (last_ == 0) { // skip to the next header
		while ((c = getc_()) >= 0 && c != '>' && c != '@') {}
		if (c < 0) return 0;
		last_ = c;
	}
	std::string hdr;
	readline_(hdr, false);
	// kseq.h: the name ends at the first isspace() character; exactly that one delimiter is consumed and the rest of the line is the comment
	size_t sp = 0;
	while (sp < hdr.size() && !isspace((unsigned char)hdr[sp])) ++sp;
	if (sp == hdr.size()) sp = std::string::npos;
	r.name = sp == std::string::npos? hdr : hdr.substr(0, sp);
	r.comment.clear();
	if (with_comment && sp != std::string::npos && sp + 1 < hdr.size()) r.comment = hdr.substr(sp + 1);
	r.seq.clear(); r.qual.clear();
	const bool is_fq = last_ == '@';
	last_ = 0;
	std::string line;
	while ((c = peek_()) >= 0 && c != '>' && c != '+' && c != '@') {
		readline_(line, false);
		// kseq keeps graph characters only; lines are almost always clean, so test first (vectorisable) and append in one go
		bool clean = true;
		for (char ch : line) clean &= ch > ' ';
		if (clean) r.seq.append(line);
		else for (char ch : line) if (ch > ' ') r.seq.push_back(ch);
	}
	{
		bool has_u = false;
		for (char ch : r.seq) has_u |= (ch == 'u') | (ch == 'U');
		if (has_u) for (char &ch : r.seq) if (ch == 'u' || ch == 'U') --ch;
	}
	if (c == '>' || c == '@') { getc_(); last_ = c; }
	if (c != '+') return 1;
	// quality block
	readline_(line, false); // the '+' line
	while (r.qual.size() < r.seq.size() && peek_() >= 0) {
		readline_(line, false);
		for (char ch : line) if (ch > ' ') r.qual.push_back(ch);
	}
	last_ = 0;
	if (r.qual.size() != r.seq.size()) return -2;
	if (!with_qual || !is_fq) r.qual.clear();
	return 1;
}
