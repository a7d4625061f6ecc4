// minimap2_b200/csrc/fastx.h -- gz-capable FASTA/FASTQ record reader (host I/O; same record semantics as the
// reference's kseq.h-based bseq.c:80-119: name = header up to the first blank, comment = the rest, multi-line
// sequences, U->T on input (bseq.c:72-74)).
#pragma once
#include <string>
#include <vector>
#include <cstdint>
#include <zlib.h>

struct FastxRecord {
	std::string name, comment, seq, qual;
};

class FastxReader {
public:
	explicit FastxReader(const char *fn);
	~FastxReader();
	bool ok() const { return fp_ != nullptr || fd_ >= 0; }
	// returns 1 on success, 0 at EOF, <0 on a malformed record
	int next(FastxRecord &r, bool with_qual, bool with_comment);
	bool eof();
private:
	int getc_();
	int peek_();
	void readline_(std::string &s, bool append);
	void seqline_(std::string &seq);   // one sequence line appended straight from the read buffer (graph characters only, U->T)
	int fill_();                       // next buffer-full: plain read(2) for uncompressed input, zlib otherwise
	gzFile fp_ = nullptr;
	int fd_ = -1;                      // uncompressed input is read without the zlib layer (one copy less per byte)
	size_t seq_hint_ = 0;              // length of the previous record: the next one reserves it up front
	std::vector<unsigned char> buf_;
	size_t beg_ = 0, end_ = 0;
	bool eof_ = false;
	int last_ = 0; // pending header char ('>' or '@') already consumed
};
