// mmio.cu -- Matrix Market ingestion (host code): the reference's benchmark scripts load their real-world operators with
// MatrixMarket.jl / MAT.jl (benchmark/matrixmarket.jl:2,9-10, benchmark/matrixcollection.jl) -- SURVEY.md section 8(f)
// item 3.  Reads the `coordinate` format (real / integer / pattern; general / symmetric / skew-symmetric) into the
// arrays of a SparseMatrixCSC{Float64,Int64}: columns ascending, rows ascending inside a column, duplicates summed
// (what `sparse(I, J, V, m, n)` does in MatrixMarket.jl's mmread), symmetric storage expanded.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

using namespace b200;

namespace {

struct MmHeader {
  int64_t m = 0, n = 0, entries = 0;
  int field = 0;      // 0 real, 1 integer, 2 pattern
  int symmetry = 0;   // 0 general, 1 symmetric, 2 skew-symmetric
  long data_pos = 0;
};

int parse_header(FILE *f, const char *path, MmHeader *h) {
  char line[1024];
  B200_REQUIRE(fgets(line, sizeof(line), f), "%s: empty file", path);
  char banner[64], object[64], format[64], field[64], symmetry[64];
  B200_REQUIRE(sscanf(line, "%63s %63s %63s %63s %63s", banner, object, format, field, symmetry) == 5 &&
                   strcmp(banner, "%%MatrixMarket") == 0,
               "%s: not a Matrix Market file (bad banner)", path);
  auto lower = [](char *s) { for (; *s; ++s) *s = (char)tolower((unsigned char)*s); };
  lower(object); lower(format); lower(field); lower(symmetry);
  B200_REQUIRE(strcmp(object, "matrix") == 0, "%s: object `%s` is not supported", path, object);
  B200_REQUIRE(strcmp(format, "coordinate") == 0, "%s: only the coordinate (sparse) format is supported, got `%s`", path,
               format);
  if (strcmp(field, "real") == 0 || strcmp(field, "double") == 0) h->field = 0;
  else if (strcmp(field, "integer") == 0) h->field = 1;
  else if (strcmp(field, "pattern") == 0) h->field = 2;
  else B200_REQUIRE(false, "%s: field `%s` is not supported (real, integer, pattern)", path, field);
  if (strcmp(symmetry, "general") == 0) h->symmetry = 0;
  else if (strcmp(symmetry, "symmetric") == 0) h->symmetry = 1;
  else if (strcmp(symmetry, "skew-symmetric") == 0) h->symmetry = 2;
  else B200_REQUIRE(false, "%s: symmetry `%s` is not supported", path, symmetry);
  for (;;) {                                           // comments and blank lines
    B200_REQUIRE(fgets(line, sizeof(line), f), "%s: missing size line", path);
    const char *p = line;
    while (*p == ' ' || *p == '\t') ++p;
    if (*p == '%' || *p == '\n' || *p == '\r' || *p == 0) continue;
    long long m, n, e;
    B200_REQUIRE(sscanf(p, "%lld %lld %lld", &m, &n, &e) == 3 && m >= 0 && n >= 0 && e >= 0, "%s: bad size line", path);
    h->m = m; h->n = n; h->entries = e;
    break;
  }
  B200_REQUIRE(h->symmetry == 0 || h->m == h->n, "%s: symmetric storage needs a square matrix", path);
  h->data_pos = ftell(f);
  return B200_OK;
}

struct Entry {
  int64_t col, row;
  double val;
};

int read_entries(FILE *f, const char *path, const MmHeader &h, std::vector<Entry> *out) {
  out->clear();
  out->reserve((size_t)(h.symmetry ? 2 * h.entries : h.entries));
  char line[1024];
  for (int64_t k = 0; k < h.entries; ++k) {
    B200_REQUIRE(fgets(line, sizeof(line), f), "%s: %lld entries announced, %lld found", path, (long long)h.entries,
                 (long long)k);
    char *p = line, *end;
    const long long i = strtoll(p, &end, 10);
    B200_REQUIRE(end != p, "%s: bad entry line %lld", path, (long long)k + 1);
    p = end;
    const long long j = strtoll(p, &end, 10);
    B200_REQUIRE(end != p, "%s: bad entry line %lld", path, (long long)k + 1);
    p = end;
    double v = 1.0;
    if (h.field != 2) {
      v = strtod(p, &end);
      B200_REQUIRE(end != p, "%s: entry line %lld has no value", path, (long long)k + 1);
    }
    B200_REQUIRE(i >= 1 && i <= h.m && j >= 1 && j <= h.n, "%s: entry (%lld, %lld) outside %lld x %lld", path, i, j,
                 (long long)h.m, (long long)h.n);
    out->push_back({j - 1, i - 1, v});
    if (h.symmetry && i != j) out->push_back({i - 1, j - 1, h.symmetry == 2 ? -v : v});
  }
  std::stable_sort(out->begin(), out->end(),
                   [](const Entry &a, const Entry &b) { return a.col != b.col ? a.col < b.col : a.row < b.row; });
  size_t w = 0;                                        // sum duplicates (sparse(I, J, V) semantics)
  for (size_t r = 0; r < out->size(); ++r) {
    if (w > 0 && (*out)[w - 1].col == (*out)[r].col && (*out)[w - 1].row == (*out)[r].row) (*out)[w - 1].val += (*out)[r].val;
    else (*out)[w++] = (*out)[r];
  }
  out->resize(w);
  return B200_OK;
}

}  // namespace

extern "C" {

int b200_mm_info(const char *path, int64_t *m, int64_t *n, int64_t *nnz, int *field, int *symmetry) {
  B200_REQUIRE(path, "NULL path");
  FILE *f = fopen(path, "r");
  B200_REQUIRE(f, "cannot open %s", path);
  MmHeader h;
  int st = parse_header(f, path, &h);
  std::vector<Entry> e;
  if (st == B200_OK) st = read_entries(f, path, h, &e);
  fclose(f);
  if (st != B200_OK) return st;
  if (m) *m = h.m;
  if (n) *n = h.n;
  if (nnz) *nnz = (int64_t)e.size();
  if (field) *field = h.field;
  if (symmetry) *symmetry = h.symmetry;
  return B200_OK;
}

int b200_mm_read_csc_i64(const char *path, int base, int64_t nnz_capacity, int64_t *colptr, int64_t *rowval,
                         double *nzval) {
  B200_REQUIRE(path && colptr && (nnz_capacity == 0 || (rowval && nzval)), "NULL argument");
  FILE *f = fopen(path, "r");
  B200_REQUIRE(f, "cannot open %s", path);
  MmHeader h;
  int st = parse_header(f, path, &h);
  std::vector<Entry> e;
  if (st == B200_OK) st = read_entries(f, path, h, &e);
  fclose(f);
  if (st != B200_OK) return st;
  B200_REQUIRE((int64_t)e.size() <= nnz_capacity, "%s holds %lld nonzeros, the arrays %lld", path, (long long)e.size(),
               (long long)nnz_capacity);
  for (int64_t j = 0; j <= h.n; ++j) colptr[j] = 0;
  for (const Entry &x : e) colptr[x.col + 1] += 1;
  colptr[0] = base;
  for (int64_t j = 0; j < h.n; ++j) colptr[j + 1] += colptr[j];
  for (size_t k = 0; k < e.size(); ++k) {
    rowval[k] = e[k].row + base;
    nzval[k] = e[k].val;
  }
  return B200_OK;
}

}  // extern "C"
