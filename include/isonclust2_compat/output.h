// output.h belongs to the reference's `dump` subcommand (src/output.h), outside the hot path; the header exists so that
// translation units written against the reference's include list compile.
#ifndef OUTPUT_H_INCLUDED
#define OUTPUT_H_INCLUDED
#endif
