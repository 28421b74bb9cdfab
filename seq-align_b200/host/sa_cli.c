/*
 * sa_cli.c -- option parsing, sequence reading and scoring-file loading: the
 * host front-end of the library, used by the batching command-line tools and,
 * through sa_cmdline.c, by callers of the reference's cmdline_* / align_from_file
 * / align_scoring_load_* functions.  See sa_cli.h for the reference code each
 * part stands in for; the flag set, defaults and error conditions follow
 * reference src/alignment_cmdline.c:179-485.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>
#include <zlib.h>

#include "sa_cli.h"

/* ---- usage / errors ------------------------------------------------------ */

static const char *g_prog = "seq-align";
static int g_tool = SA_TOOL_NW;
static int g_defaults[4];

static void die_usage(const char *fmt, ...) __attribute__((format(printf, 1, 2), noreturn));

static void die_usage(const char *fmt, ...)
{
  const int sw = g_tool == SA_TOOL_SW;
  if(fmt) {
    va_list ap;
    fputs("Error: ", stderr);
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    if(fmt[0] == '\0' || fmt[strlen(fmt) - 1] != '\n') fputc('\n', stderr);
  }
  fprintf(stderr, "usage: %s [OPTIONS] [seq1 seq2]\n", g_prog);
  fprintf(stderr,
          "  %s optimal %s alignment (maximises score), B200 batch engine.\n"
          "  Takes a pair of sequences on the command line, or reads pairs from\n"
          "  files / stdin (FASTA, FASTQ or one sequence per line; gzip ok).\n"
          "  Pairs read from files are aligned in batches on the GPU.\n\n",
          sw ? "Smith-Waterman" : "Needleman-Wunsch", sw ? "local" : "global");
  fprintf(stderr,
          "  OPTIONS:\n"
          "    --file <file>        read two sequences at a time from <file> and align them\n"
          "    --files <f1> <f2>    read one sequence from each file at a time\n"
          "    --stdin              read from STDIN (same as '--file -'), answer pair by pair\n"
          "    --gpus <n|all>       cut the batches over <n> GPUs of this machine [default: 1]\n\n"
          "    --case_sensitive     case sensitive character comparison [default: off]\n\n"
          "    --match <score>      [default: %i]\n"
          "    --mismatch <score>   [default: %i]\n"
          "    --gapopen <score>    [default: %i]\n"
          "    --gapextend <score>  [default: %i]\n\n"
          "    --scoring <PAM30|PAM70|BLOSUM80|BLOSUM62>\n"
          "    --substitution_matrix <file>\n"
          "    --substitution_pairs <file>\n\n"
          "    --wildcard <w> <s>   character <w> matches all characters with score <s>\n\n",
          g_defaults[0], g_defaults[1], g_defaults[2], g_defaults[3]);
  if(sw)
    fprintf(stderr,
            "    --minscore <score>   minimum required score [default: match * MAX(0.2 * length, 2)]\n"
            "    --maxhits <hits>     maximum number of results per alignment [default: no limit]\n\n"
            "    --context <n>        print <n> bases of context\n"
            "    --printseq           print sequences before local alignments\n");
  else
    fprintf(stderr,
            "\n"
            "    --freestartgap       no penalty for gap at start of alignment\n"
            "    --freeendgap         no penalty for gap at end of alignment\n\n"
            "    --printscores        print optimal alignment scores\n"
            "    --zam                a funky type of output\n");
  fprintf(stderr,
          "    --printmatrices      print dynamic programming matrices\n"
          "    --printfasta         print fasta header lines\n"
          "    --pretty             print with a descriptor line\n"
          "    --colour             print with colour\n\n"
          "  Experimental options:\n"
          "    --nogapsin1          no gaps allowed within the first sequence\n"
          "    --nogapsin2          no gaps allowed within the second sequence\n"
          "    --nogaps             no gaps allowed in either sequence\n"
          "    --nomismatches       no mismatches allowed%s\n\n"
          "  Gap (of length N) penalty is: (open+N*extend); '--gapopen 0' gives linear gaps.\n",
          sw ? "" : " (cannot be used with --nogaps..)");
  exit(EXIT_FAILURE);
}

static int whole_int(const char *s, int *out)
{
  char *end = NULL;
  const long v = strtol(s, &end, 10);
  if(v > INT_MAX || v < INT_MIN || end != s + strlen(s)) return 0;
  *out = (int)v;
  return 1;
}

static int whole_uint(const char *s, unsigned *out)
{
  char *end = NULL;
  const unsigned long v = strtoul(s, &end, 10);
  if(v > UINT_MAX || end != s + strlen(s)) return 0;
  *out = (unsigned)v;
  return 1;
}

/* ---- options ------------------------------------------------------------- */

static void add_files(sa_opts *o, const char *p1, const char *p2)
{
  if(o->nfiles == o->files_cap) {
    o->files_cap = o->files_cap ? 2 * o->files_cap : 16;
    o->files = realloc(o->files, o->files_cap * sizeof(sa_file_pair));
    if(!o->files) { fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__); exit(EXIT_FAILURE); }
  }
  o->files[o->nfiles].path1 = p1;
  o->files[o->nfiles].path2 = p2;
  o->nfiles++;
}

void sa_cli_free(sa_opts *o)
{
  free(o->files);
  o->files = NULL;
  o->nfiles = o->files_cap = 0;
}

/* kinds of option: which tool may use it and how many parameters follow */
enum { ANY = 0, NW_ONLY = 1, SW_ONLY = 2 };
enum {
  O_FREESTART, O_FREEEND, O_NOGAPS, O_NOGAPS1, O_NOGAPS2, O_NOMISMATCH, O_CASE, O_PRINTSEQ, O_PRINTMAT,
  O_PRINTSCORES, O_PRINTFASTA, O_PRETTY, O_COLOUR, O_ZAM, O_STDIN,
  /* one parameter */
  O_SCORING, O_SUBMATRIX, O_SUBPAIRS, O_MINSCORE, O_MAXHITS, O_CONTEXT, O_MATCH, O_MISMATCH, O_GAPOPEN,
  O_GAPEXTEND, O_FILE, O_GPUS,
  /* two parameters, checked by the option itself */
  O_FILES, O_WILDCARD
};
static const struct { const char *name; int id, nparam, who; const char *only_msg; } kOptions[] = {
    {"--freestartgap", O_FREESTART, 0, NW_ONLY, "--freestartgap only valid with Needleman-Wunsch"},
    {"--freeendgap", O_FREEEND, 0, NW_ONLY, "--freeendgap only valid with Needleman-Wunsch"},
    {"--nogaps", O_NOGAPS, 0, ANY, NULL},
    {"--nogapsin1", O_NOGAPS1, 0, ANY, NULL},
    {"--nogapsin2", O_NOGAPS2, 0, ANY, NULL},
    {"--nomismatches", O_NOMISMATCH, 0, ANY, NULL},
    {"--case_sensitive", O_CASE, 0, ANY, NULL},
    {"--printseq", O_PRINTSEQ, 0, SW_ONLY, "--printseq only valid with Smith-Waterman"},
    {"--printmatrices", O_PRINTMAT, 0, ANY, NULL},
    {"--printscores", O_PRINTSCORES, 0, NW_ONLY, "--printscores only valid with Needleman-Wunsch"},
    {"--printfasta", O_PRINTFASTA, 0, ANY, NULL},
    {"--pretty", O_PRETTY, 0, ANY, NULL},
    {"--colour", O_COLOUR, 0, ANY, NULL},
    {"--zam", O_ZAM, 0, NW_ONLY, "--zam only valid with Needleman-Wunsch"},
    {"--stdin", O_STDIN, 0, ANY, NULL},
    {"--scoring", O_SCORING, 1, ANY, NULL},
    {"--substitution_matrix", O_SUBMATRIX, 1, ANY, NULL},
    {"--substitution_pairs", O_SUBPAIRS, 1, ANY, NULL},
    {"--minscore", O_MINSCORE, 1, SW_ONLY, "--minscore only valid with Smith-Waterman"},
    {"--maxhits", O_MAXHITS, 1, SW_ONLY, "--maxhits only valid with Smith-Waterman"},
    {"--context", O_CONTEXT, 1, SW_ONLY, "--context only valid with Smith-Waterman"},
    {"--match", O_MATCH, 1, ANY, NULL},
    {"--mismatch", O_MISMATCH, 1, ANY, NULL},
    {"--gapopen", O_GAPOPEN, 1, ANY, NULL},
    {"--gapextend", O_GAPEXTEND, 1, ANY, NULL},
    {"--file", O_FILE, 1, ANY, NULL},
    {"--gpus", O_GPUS, 1, ANY, NULL},
    {"--files", O_FILES, 2, ANY, NULL},
    {"--wildcard", O_WILDCARD, 2, ANY, NULL},
};

static int find_option(const char *arg)
{
  for(size_t i = 0; i < sizeof(kOptions) / sizeof(kOptions[0]); i++)
    if(strcasecmp(arg, kOptions[i].name) == 0) return (int)i;
  return -1;
}

void sa_cli_parse(int argc, char **argv, scoring_t *sc, int tool, sa_opts *o)
{
  memset(o, 0, sizeof(*o));
  o->tool = tool;
  g_tool = tool;
  g_prog = argv[0];
  g_defaults[0] = sc->match; g_defaults[1] = sc->mismatch;
  g_defaults[2] = sc->gap_open; g_defaults[3] = sc->gap_extend;
  if(argc == 1) die_usage(NULL);

  /* first sweep: help, case sensitivity (the loaders need it) and the named
   * scoring system, which later options then modify */
  int scoring_named = 0;
  for(int i = 1; i < argc; i++) {
    const char *a = argv[i];
    if(!strcasecmp(a, "--help") || !strcasecmp(a, "-help") || !strcasecmp(a, "-h")) die_usage(NULL);
    if(!strcasecmp(a, "--case_sensitive")) o->case_sensitive = 1;
    else if(!strcasecmp(a, "--scoring")) {
      if(scoring_named) die_usage("More than one scoring system specified - not permitted");
      const char *name = i + 1 < argc ? argv[i + 1] : "";
      if(!strcasecmp(name, "PAM30")) scoring_system_PAM30(sc);
      else if(!strcasecmp(name, "PAM70")) scoring_system_PAM70(sc);
      else if(!strcasecmp(name, "BLOSUM80")) scoring_system_BLOSUM80(sc);
      else if(!strcasecmp(name, "BLOSUM62")) scoring_system_BLOSUM62(sc);
      else if(!strcasecmp(name, "DNA_HYBRIDIZATION")) scoring_system_DNA_hybridization(sc);
      else die_usage("Unknown --scoring choice, not one of PAM30|PAM70|BLOSUM80|BLOSUM62");
      scoring_named = 1;
      i++;
    }
  }

  int tables_loaded = 0, match_set = 0, mismatch_set = 0;
  int i;
  for(i = 1; i < argc; i++) {
    const char *a = argv[i];
    if(a[0] != '-') {
      if(argc - i != 2) die_usage("Unknown options: '%s'", a);
      break;
    }
    const int k = find_option(a);
    if(k >= 0 && kOptions[k].nparam == 0) {
      if((kOptions[k].who == NW_ONLY && tool != SA_TOOL_NW) || (kOptions[k].who == SW_ONLY && tool != SA_TOOL_SW))
        die_usage("%s", kOptions[k].only_msg);
      switch(kOptions[k].id) {
        case O_FREESTART: sc->no_start_gap_penalty = true; break;
        case O_FREEEND: sc->no_end_gap_penalty = true; break;
        case O_NOGAPS: sc->no_gaps_in_a = sc->no_gaps_in_b = true; break;
        case O_NOGAPS1: sc->no_gaps_in_a = true; break;
        case O_NOGAPS2: sc->no_gaps_in_b = true; break;
        case O_NOMISMATCH: sc->no_mismatches = true; break;
        case O_CASE: break; /* first sweep */
        case O_PRINTSEQ: o->print_seq = 1; break;
        case O_PRINTMAT: o->print_matrices = 1; break;
        case O_PRINTSCORES: o->print_scores = 1; break;
        case O_PRINTFASTA: o->print_fasta = 1; break;
        case O_PRETTY: o->print_pretty = 1; break;
        case O_COLOUR: o->print_colour = 1; break;
        case O_ZAM: o->zam = 1; break;
        case O_STDIN: add_files(o, "", NULL); o->interactive = 1; break;
      }
      continue;
    }
    /* everything else wants at least one parameter */
    if(i == argc - 1) die_usage("Unknown argument without parameter: %s", a);
    if(k < 0) die_usage("Unknown argument '%s'", a);
    if((kOptions[k].who == NW_ONLY && tool != SA_TOOL_NW) || (kOptions[k].who == SW_ONLY && tool != SA_TOOL_SW))
      die_usage("%s", kOptions[k].only_msg);
    const char *p = argv[i + 1];
    switch(kOptions[k].id) {
      case O_SCORING: break; /* first sweep */
      case O_SUBMATRIX: sa_load_matrix(p, sc, o->case_sensitive); tables_loaded = 1; break;
      case O_SUBPAIRS: sa_load_pairs(p, sc, o->case_sensitive); tables_loaded = 1; break;
      case O_MINSCORE:
        if(!whole_int(p, &o->min_score)) die_usage("Invalid --minscore <score> argument (must be a +ve int)");
        o->min_score_set = 1;
        break;
      case O_MAXHITS:
        if(!whole_uint(p, &o->max_hits)) die_usage("Invalid --maxhits <hits> argument (must be a +ve int)");
        o->max_hits_set = 1;
        break;
      case O_CONTEXT:
        if(!whole_uint(p, &o->context)) die_usage("Invalid --context <c> argument (must be >= 0)");
        break;
      case O_MATCH:
        if(!whole_int(p, &sc->match)) die_usage("Invalid --match argument ('%s') must be an int", p);
        match_set = 1;
        break;
      case O_MISMATCH:
        if(!whole_int(p, &sc->mismatch)) die_usage("Invalid --mismatch argument ('%s') must be an int", p);
        mismatch_set = 1;
        break;
      case O_GAPOPEN:
        if(!whole_int(p, &sc->gap_open)) die_usage("Invalid --gapopen argument ('%s') must be an int", p);
        break;
      case O_GAPEXTEND:
        if(!whole_int(p, &sc->gap_extend)) die_usage("Invalid --gapextend argument ('%s') must be an int", p);
        break;
      case O_FILE: add_files(o, p, NULL); break;
      case O_GPUS:
        if(!strcasecmp(p, "all")) o->gpus = 0;
        else if(!whole_int(p, &o->gpus) || o->gpus < 1) die_usage("Invalid --gpus <n> argument (a number >= 1, or 'all')");
        o->gpus_set = 1;
        break;
      case O_FILES:
        if(i >= argc - 2) die_usage("--files option takes 2 arguments");
        if(!strcmp(p, "-") && !strcmp(argv[i + 2], "-")) add_files(o, p, NULL); /* both from stdin */
        else add_files(o, p, argv[i + 2]);
        break;
      case O_WILDCARD: {
        int ws = 0;
        if(i == argc - 2 || strlen(p) != 1 || !whole_int(argv[i + 2], &ws))
          die_usage("--wildcard <w> <s> takes a single character and a number");
        scoring_add_wildcard(sc, p[0], ws);
        break;
      }
    }
    i += kOptions[k].nparam;
  }

  if((match_set && !mismatch_set && !sc->no_mismatches) || (!match_set && mismatch_set))
    die_usage("--match --mismatch must both be set or neither set");
  else if(tables_loaded && !match_set)
    sc->use_match_mismatch = 0; /* a loaded table replaces match/mismatch */
  if(sc->use_match_mismatch && sc->match < sc->mismatch)
    die_usage("Match value should not be less than mismatch penalty");
  if(tool == SA_TOOL_NW && sc->no_mismatches && (sc->no_gaps_in_a || sc->no_gaps_in_b))
    die_usage("--nogaps.. --nomismatches cannot be used at together");
  if(i < argc) { o->seq1 = argv[i]; o->seq2 = argv[i + 1]; }
  if(!o->seq1 && o->nfiles == 0) die_usage("No input specified");
  if(o->zam && (o->print_pretty || o->print_scores || o->print_colour || o->print_fasta))
    die_usage("Cannot use --printscore, --printfasta, --pretty or --colour with --zam");
}

/* ---- growable string ----------------------------------------------------- */

static void str_reserve(sa_str *s, size_t need)
{
  if(need + 1 <= s->cap) return;
  size_t cap = s->cap ? s->cap : 256;
  while(cap < need + 1) cap *= 2;
  s->b = realloc(s->b, cap);
  if(!s->b) { fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__); exit(EXIT_FAILURE); }
  s->cap = cap;
}

static void str_clear(sa_str *s) { str_reserve(s, 0); s->len = 0; s->b[0] = '\0'; }
static void str_push(sa_str *s, int c) { str_reserve(s, s->len + 1); s->b[s->len++] = (char)c; s->b[s->len] = '\0'; }
static void str_append(sa_str *s, const char *p, size_t n)
{
  str_reserve(s, s->len + n);
  memcpy(s->b + s->len, p, n);
  s->len += n;
  s->b[s->len] = '\0';
}
static void str_chomp(sa_str *s)
{
  while(s->len && (s->b[s->len - 1] == '\n' || s->b[s->len - 1] == '\r')) s->len--;
  if(s->b) s->b[s->len] = '\0';
}
void sa_str_free(sa_str *s) { free(s->b); s->b = NULL; s->len = s->cap = 0; }

/* ---- sequence reader ------------------------------------------------------
 * Record grammar as read by the reference (seq_file.h:245-325):
 *   FASTA  '>'name  then lines up to the next line starting with '>'
 *   FASTQ  '@'name  sequence lines up to a line starting with '+', then
 *          quality lines until at least as many characters as bases
 *   plain  one sequence per line; lines starting with white space are skipped
 * The format is decided by the first non-blank character of each record. */

enum { FMT_UNKNOWN = 0, FMT_FASTA, FMT_FASTQ, FMT_PLAIN };

struct sa_reader {
  gzFile gz;
  int fd;          /* unbuffered path (interactive stdin) */
  int fmt;
  int pushed;      /* one character of push-back, -2 = none */
  unsigned char *buf;
  size_t pos, end, cap;
  int eof;
  sa_str qual;
};

sa_reader *sa_reader_open(const char *path, int buffered)
{
  sa_reader *r = calloc(1, sizeof(*r));
  if(!r) return NULL;
  r->pushed = -2;
  r->fd = -1;
  if(!buffered && !strcmp(path, "-")) {
    r->fd = STDIN_FILENO;
    return r;
  }
  r->gz = !strcmp(path, "-") ? gzdopen(dup(STDIN_FILENO), "r") : gzopen(path, "r");
  if(!r->gz) { free(r); return NULL; }
  r->cap = 1 << 20;
  r->buf = malloc(r->cap);
  gzbuffer(r->gz, 1 << 18);
  return r;
}

void sa_reader_close(sa_reader *r)
{
  if(!r) return;
  if(r->gz) gzclose(r->gz);
  free(r->buf);
  sa_str_free(&r->qual);
  free(r);
}

static int rd_fill(sa_reader *r)
{
  if(r->eof) return 0;
  if(r->gz) {
    const int n = gzread(r->gz, r->buf, (unsigned)r->cap);
    if(n <= 0) { r->eof = 1; return 0; }
    r->pos = 0; r->end = (size_t)n;
    return 1;
  }
  return 0;
}

static int rd_getc(sa_reader *r)
{
  if(r->pushed != -2) { const int c = r->pushed; r->pushed = -2; return c; }
  if(r->fd >= 0) {
    unsigned char c;
    if(r->eof) return -1;
    const ssize_t n = read(r->fd, &c, 1);
    if(n <= 0) { r->eof = 1; return -1; }
    return c;
  }
  if(r->pos == r->end && !rd_fill(r)) return -1;
  return r->buf[r->pos++];
}

static void rd_ungetc(sa_reader *r, int c) { r->pushed = c; }

int sa_reader_getc(sa_reader *r) { return rd_getc(r); }

/* append the rest of the current line (newline included) to s; number of characters read */
static size_t rd_line(sa_reader *r, sa_str *s)
{
  size_t got = 0;
  if(r->pushed != -2) {
    const int c = rd_getc(r);
    if(c == -1) return 0;
    str_push(s, c); got++;
    if(c == '\n') return got;
  }
  if(r->fd >= 0) {
    int c;
    while((c = rd_getc(r)) != -1) { str_push(s, c); got++; if(c == '\n') break; }
    return got;
  }
  for(;;) {
    if(r->pos == r->end && !rd_fill(r)) return got;
    const unsigned char *p = r->buf + r->pos;
    const size_t avail = r->end - r->pos;
    const unsigned char *nl = memchr(p, '\n', avail);
    const size_t take = nl ? (size_t)(nl - p) + 1 : avail;
    str_append(s, (const char *)p, take);
    r->pos += take; got += take;
    if(nl) return got;
  }
}

/* A line that starts with white space: the reference means to skip it, and does when it reads
 * stdin unbuffered (--stdin); through its buffered reader (--file, --files) the skip has no effect
 * and only the leading white space goes -- the rest of the line is read as a sequence (observed on
 * the reference's tools; seq_file.h:303,316 call a skipline that does not move the buffer).  Both
 * behaviours are kept, each where the reference shows it. */
static void rd_skipline(sa_reader *r)
{
  if(r->fd < 0) return;
  int c;
  while((c = rd_getc(r)) != -1 && c != '\n') {}
}

static int read_fasta(sa_reader *r, sa_record *rec)
{
  int c = rd_getc(r);
  if(c == -1) return 0;
  if(c != '>' || rd_line(r, &rec->name) == 0) return -1;
  str_chomp(&rec->name);
  while((c = rd_getc(r)) != '>') {
    if(c == -1) return 1;
    if(c != '\r' && c != '\n') {
      str_push(&rec->seq, c);
      const size_t n = rd_line(r, &rec->seq);
      str_chomp(&rec->seq);
      if(n == 0) return 1;
    }
  }
  rd_ungetc(r, c);
  return 1;
}

static int read_fastq(sa_reader *r, sa_record *rec)
{
  int c = rd_getc(r);
  if(c == -1) return 0;
  if(c != '@' || rd_line(r, &rec->name) == 0) return -1;
  str_chomp(&rec->name);
  while((c = rd_getc(r)) != '+') {
    if(c == -1) return -1;
    if(c != '\r' && c != '\n') {
      str_push(&rec->seq, c);
      if(rd_line(r, &rec->seq) == 0) return -1;
      str_chomp(&rec->seq);
    }
  }
  while((c = rd_getc(r)) != -1 && c != '\n') {}
  if(c == -1) return -1;
  str_clear(&r->qual);
  do {
    if(rd_line(r, &r->qual) > 0) str_chomp(&r->qual);
    else return 1;
  } while(r->qual.len < rec->seq.len);
  while((c = rd_getc(r)) != -1 && c != '@') {}
  if(c != -1) rd_ungetc(r, c);
  return 1;
}

static int read_plain(sa_reader *r, sa_record *rec)
{
  int c;
  while((c = rd_getc(r)) != -1 && isspace(c)) if(c != '\n') rd_skipline(r);
  if(c == -1) return 0;
  str_push(&rec->seq, c);
  rd_line(r, &rec->seq);
  str_chomp(&rec->seq);
  return 1;
}

int sa_reader_next(sa_reader *r, sa_record *rec)
{
  str_clear(&rec->name);
  str_clear(&rec->seq);
  /* The format is chosen per RECORD, from its first non-blank character: the reference's seq_read()
   * goes through the "unknown format" reader every time (seq_file.h:97 calls readfunc, which
   * :318-320 never re-point), so a line starting with '>' or '@' inside a one-per-line file opens
   * a FASTA / FASTQ record there. */
  {
    int c;
    while((c = rd_getc(r)) != -1 && isspace(c)) if(c != '\n') rd_skipline(r);
    if(c == -1) return 0;
    r->fmt = c == '@' ? FMT_FASTQ : c == '>' ? FMT_FASTA : FMT_PLAIN;
    rd_ungetc(r, c);
  }
  switch(r->fmt) {
    case FMT_FASTA: return read_fasta(r, rec);
    case FMT_FASTQ: return read_fastq(r, rec);
    default: return read_plain(r, rec);
  }
}

/* ---- scoring files --------------------------------------------------------
 * Formats of reference src/alignment_scoring_load.c:39-306. */

static void load_error(int matrix, const char *msg, const char *path, int with_line) __attribute__((noreturn));
static void load_error(int matrix, const char *msg, const char *path, int with_line)
{
  fprintf(stderr, matrix ? "Error: substitution matrix : %s\n" : "Error: substitution pairs : %s\n", msg);
  if(path) fprintf(stderr, "File: %s\n", path);
  if(with_line) fprintf(stderr, "Line: %s\n", path);
  exit(EXIT_FAILURE);
}

static int gz_line(gzFile f, sa_str *s)
{
  char tmp[4096];
  str_clear(s);
  while(gzgets(f, tmp, (int)sizeof(tmp))) {
    const size_t n = strlen(tmp);
    str_append(s, tmp, n);
    if(n && tmp[n - 1] == '\n') break;
  }
  return s->len > 0;
}

static int all_space(const char *p)
{
  for(; *p; p++) if(!isspace((unsigned char)*p)) return 0;
  return 1;
}

static char *next_nonspace(char *p)
{
  while(*p && isspace((unsigned char)*p)) p++;
  return *p ? p : NULL;
}

static gzFile open_scoring_file(const char *path)
{
  gzFile f = gzopen(path, "r");
  if(!f) die_usage("Couldn't read: %s", path);
  return f;
}

#define FOLD(c) (case_sensitive ? (char)(c) : (char)tolower((unsigned char)(c)))

void sa_load_matrix(const char *path, scoring_t *sc, int case_sensitive)
{
  gzFile f = open_scoring_file(path);
  sa_load_matrix_gz(f, path, sc, case_sensitive);
  gzclose(f);
}

void sa_load_pairs(const char *path, scoring_t *sc, int case_sensitive)
{
  gzFile f = open_scoring_file(path);
  sa_load_pairs_gz(f, path, sc, case_sensitive);
  gzclose(f);
}

void sa_load_matrix_gz(void *gz, const char *path, scoring_t *sc, int case_sensitive)
{
  gzFile f = (gzFile)gz;
  sa_str ln = {0, 0, 0};
  int lines = 0, have_header = 0;
  /* heading row: first line that is not blank and not a comment */
  while(gz_line(f, &ln)) {
    str_chomp(&ln);
    if(ln.len > 0 && ln.b[0] != '#' && !all_space(ln.b)) {
      if(ln.len < 2) load_error(1, "Too few column headings", path, 1);
      have_header = 1;
      break;
    }
    lines++;
  }
  if(!have_header && lines == 0) load_error(1, "Empty file", path, 0);
  if(!have_header) { sa_str_free(&ln); return; }
  const char sep = ln.b[0];
  if((sep >= '0' && sep <= '9') || sep == '-')
    load_error(1, "Numbers (0-9) and dashes (-) do not make good separators", path, 0);
  char *cols = malloc(ln.len + 1);
  int ncols = 0;

  if(isspace((unsigned char)sep)) {
    /* white-space separated (the NCBI layout): every visible character of the heading is a column */
    for(char *p = ln.b; (p = next_nonspace(p + 1)) != NULL;) cols[ncols++] = FOLD(*p);
    while(gz_line(f, &ln)) {
      str_chomp(&ln);
      char *first = next_nonspace(ln.b);
      if(!first || ln.b[0] == '#') continue;
      const char from = FOLD(*first);
      char *p = ln.b + 1;
      for(int c = 0; c < ncols; c++) {
        if(!isspace((unsigned char)*p)) load_error(1, "Expected whitespace between elements - found character", path, 1);
        p = next_nonspace(p + 1);
        char *end = p;
        const int v = p ? (int)strtol(p, &end, 10) : 0;
        if(!p || end == p) load_error(1, "Missing number value on line", path, 1);
        scoring_add_mutation(sc, from, cols[c], v);
        p = end;
      }
      if(*p != '\0' && !all_space(p)) load_error(1, "Too many columns on row", path, 1);
    }
  } else {
    /* single-character separator: <sep>c<sep>c... heading, rows of <sep>number */
    for(size_t k = 0; k < ln.len; k += 2) {
      if(ln.b[k] != sep) load_error(1, "Separator missing from line", path, 1);
      cols[ncols++] = FOLD(ln.b[k + 1]);
    }
    while(gz_line(f, &ln)) {
      str_chomp(&ln);
      const char from = FOLD(ln.b[0]);
      if(from == '#' || all_space(ln.b)) continue;
      char *p = ln.b;
      int c = 0;
      while(*p != '\0') {
        const char to = cols[c++];
        if(*p != sep) load_error(1, "Separator missing from line", path, 1);
        p++;
        char *end = p;
        const int v = (int)strtol(p, &end, 10);
        if(end == p) load_error(1, "Missing number value on line", path, 1);
        if(c >= ncols) load_error(1, "Too many columns on row", path, 1);
        scoring_add_mutation(sc, from, to, v);
        p = end;
      }
    }
  }
  free(cols);
  sa_str_free(&ln);
}

void sa_load_pairs_gz(void *gz, const char *path, scoring_t *sc, int case_sensitive)
{
  gzFile f = (gzFile)gz;
  sa_str ln = {0, 0, 0};
  int added = 0;
  while(gz_line(f, &ln)) {
    const size_t raw = ln.len;
    str_chomp(&ln);
    if(ln.len == 0 || ln.b[0] == '#' || all_space(ln.b)) continue;
    if(raw < 5) load_error(0, "Too few column headings", path, 0);
    char a, b;
    int v = 0;
    if(isspace((unsigned char)ln.b[1])) {
      /* "a b score", white-space separated */
      size_t k = 1;
      while(ln.b[k] != '\0' && isspace((unsigned char)ln.b[k])) k++;
      if(k + 2 >= ln.len || !isspace((unsigned char)ln.b[k + 1])) load_error(0, "Line too short", path, 0);
      a = ln.b[0]; b = ln.b[k];
      if(!whole_int(ln.b + k + 2, &v)) load_error(0, "Invalid number", path, 0);
    } else {
      /* "a<sep>b<sep>score" with one separator character */
      if(ln.b[1] != ln.b[3]) load_error(0, "Inconsistent separators used", path, 0);
      a = ln.b[0]; b = ln.b[2];
      if(!whole_int(ln.b + 4, &v)) load_error(0, "Invalid number", path, 0);
    }
    scoring_add_mutation(sc, FOLD(a), FOLD(b), v);
    added++;
  }
  sa_str_free(&ln);
  if(!added) load_error(0, "No pairs added from file (file empty?)", path, 0);
}
