// minimap2_b200/csrc/fastx.cc -- see fastx.h
#include "fastx.h"
#include <cctype>
#include <cstring>
#include <unistd.h>
#include <fcntl.h>
#include <cerrno>

FastxReader::FastxReader(const char *fn)
{
	buf_.resize(4 << 20);
	if (fn && strcmp(fn, "-") != 0) { // a regular file that does not start with the gzip magic is read directly
		int fd = open(fn, O_RDONLY);
		if (fd >= 0) {
			unsigned char m[2];
			const ssize_t k = pread(fd, m, 2, 0);
			if (!(k == 2 && m[0] == 0x1f && m[1] == 0x8b)) { fd_ = fd; return; }
			close(fd);
		}
	}
	fp_ = (fn && strcmp(fn, "-") != 0)? gzopen(fn, "r") : gzdopen(0, "r");
	if (fp_) gzbuffer(fp_, 1 << 20);
}

FastxReader::~FastxReader() { if (fp_) gzclose(fp_); if (fd_ >= 0) close(fd_); }

int FastxReader::fill_()
{
	if (fd_ >= 0) {
		ssize_t n;
		do n = read(fd_, buf_.data(), buf_.size()); while (n < 0 && errno == EINTR);
		return (int)n;
	}
	return gzread(fp_, buf_.data(), (unsigned)buf_.size());
}

int FastxReader::getc_()
{
	if (beg_ >= end_) {
		if (eof_) return -1;
		int n = fill_();
		if (n <= 0) { eof_ = true; return -1; }
		beg_ = 0, end_ = (size_t)n;
	}
	return buf_[beg_++];
}

void FastxReader::seqline_(std::string &seq)
{
	for (;;) {
		if (beg_ >= end_) { if (getc_() < 0) return; --beg_; }
		const unsigned char *p = buf_.data() + beg_, *e = buf_.data() + end_;
		const unsigned char *nl = (const unsigned char*)memchr(p, '\n', e - p);
		const size_t n = (nl? nl : e) - p, at = seq.size();
		seq.append((const char*)p, n);
		// kseq keeps graph characters only and bseq.c:72-74 turns U into T: lines are almost always clean, so one branch-free pass decides
		unsigned bad = 0;
		const unsigned char *q = (const unsigned char*)seq.data() + at;
		for (size_t i = 0; i < n; ++i) bad |= (unsigned)(q[i] <= ' ') | (unsigned)((q[i] | 0x20) == 'u');
		if (bad) {
			size_t w = at;
			for (size_t i = at; i < at + n; ++i) { char ch = seq[i]; if (ch > ' ') { if (ch == 'u' || ch == 'U') --ch; seq[w++] = ch; } }
			seq.resize(w);
		}
		beg_ += n + (nl? 1 : 0);
		if (nl) return;
	}
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
	if (seq_hint_) r.seq.reserve(seq_hint_ + (seq_hint_ >> 3) + 64);
	while ((c = peek_()) >= 0 && c != '>' && c != '+' && c != '@') seqline_(r.seq);
	seq_hint_ = r.seq.size();
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
