/*
 * CPU restatement of the HNSW index the reference reaches through chromadb==0.5.3 ->
 * chroma-hnswlib (requirements.txt:6) -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * The dependency's source is not in /root/reference and cannot be installed here, so this is
 * a from-scratch restatement of the *published* algorithm (Malkov & Yashunin, "Efficient and
 * robust approximate nearest neighbor search using Hierarchical Navigable Small World graphs")
 * at the parameters the reference's shipped index pins (vector_store/70ef2421-.../header.bin,
 * SURVEY.md 8c): M = 16, maxM0 = 32, ef_construction = 100, level multiplier 1/ln(M), and
 * Chroma's default search_ef = 10, space l2 (squared).  It exists so that bench.py can print an
 * HNSW recall / QPS figure *labelled as a restatement* next to the exact search; it is NOT
 * Chroma, and nothing in the product links or calls it.
 *
 * Build (done by __graft_entry__.build()):  gcc -O3 -march=native -fopenmp -shared -fPIC
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int dim, M, M0, efc;
  int64_t n, cap;
  const float* data;      /* caller-owned [n][dim] */
  int* level;             /* level of each node */
  int* links0;            /* [cap][M0+1]: count then neighbours, level 0 */
  int** linksU;           /* per node: [level][M+1] for levels >= 1 */
  int maxlevel, entry;
  double mult;
  uint64_t rng;
} hnsw_t;

static float l2sqr(const float* a, const float* b, int d) {
  float s = 0.f;
  for (int i = 0; i < d; ++i) { float t = a[i] - b[i]; s += t * t; }
  return s;
}

static double urand(hnsw_t* h) {       /* xorshift64* */
  h->rng ^= h->rng >> 12; h->rng ^= h->rng << 25; h->rng ^= h->rng >> 27;
  return ((h->rng * 2685821657736338717ull) >> 11) * (1.0 / 9007199254740992.0);
}

static int* links_of(hnsw_t* h, int node, int lvl) {
  return lvl == 0 ? h->links0 + (size_t)node * (h->M0 + 1) : h->linksU[node] + (size_t)(lvl - 1) * (h->M + 1);
}

/* binary heaps of (dist, id) */
typedef struct { float d; int id; } cand_t;
typedef struct { cand_t* a; int n, cap; } heap_t;
static void heap_init(heap_t* q, int cap) { q->a = (cand_t*)malloc(sizeof(cand_t) * (size_t)cap); q->n = 0; q->cap = cap; }
static void heap_push(heap_t* q, cand_t c, int maxheap) {
  if (q->n == q->cap) { q->cap *= 2; q->a = (cand_t*)realloc(q->a, sizeof(cand_t) * (size_t)q->cap); }
  int i = q->n++;
  while (i > 0) {
    int p = (i - 1) / 2;
    int up = maxheap ? (q->a[p].d < c.d) : (q->a[p].d > c.d);
    if (!up) break;
    q->a[i] = q->a[p]; i = p;
  }
  q->a[i] = c;
}
static cand_t heap_pop(heap_t* q, int maxheap) {
  cand_t top = q->a[0], last = q->a[--q->n];
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, c = l;
    if (l >= q->n) break;
    if (r < q->n && (maxheap ? q->a[r].d > q->a[l].d : q->a[r].d < q->a[l].d)) c = r;
    int down = maxheap ? (q->a[c].d > last.d) : (q->a[c].d < last.d);
    if (!down) break;
    q->a[i] = q->a[c]; i = c;
  }
  q->a[i] = last;
  return top;
}

/* searchBaseLayer: best-first search at one level with a result set of size ef.
 * `visited` is an epoch array.  Results are left in `res` (max-heap, worst on top). */
static void search_layer(hnsw_t* h, const float* q, int ep, float epd, int lvl, int ef, heap_t* res,
                         heap_t* cands, unsigned* visited, unsigned epoch) {
  res->n = 0; cands->n = 0;
  cand_t c = {epd, ep};
  heap_push(res, c, 1); heap_push(cands, c, 0);
  visited[ep] = epoch;
  while (cands->n) {
    cand_t cur = heap_pop(cands, 0);
    if (cur.d > res->a[0].d && res->n >= ef) break;
    int* l = links_of(h, cur.id, lvl);
    for (int i = 1; i <= l[0]; ++i) {
      int nb = l[i];
      if (visited[nb] == epoch) continue;
      visited[nb] = epoch;
      float d = l2sqr(q, h->data + (size_t)nb * h->dim, h->dim);
      if (res->n < ef || d < res->a[0].d) {
        cand_t e = {d, nb};
        heap_push(cands, e, 0); heap_push(res, e, 1);
        if (res->n > ef) heap_pop(res, 1);
      }
    }
  }
}

/* getNeighborsByHeuristic2: keep a candidate only if it is closer to the base point than to
 * every neighbour kept so far.  `c` sorted ascending by distance to the base; returns count. */
static int select_heuristic(hnsw_t* h, cand_t* c, int n, int M, int* out) {
  int k = 0;
  for (int i = 0; i < n && k < M; ++i) {
    int good = 1;
    const float* ci = h->data + (size_t)c[i].id * h->dim;
    for (int j = 0; j < k; ++j) {
      if (l2sqr(ci, h->data + (size_t)out[j] * h->dim, h->dim) < c[i].d) { good = 0; break; }
    }
    if (good) out[k++] = c[i].id;
  }
  return k;
}

static int cmp_cand(const void* a, const void* b) {
  float x = ((const cand_t*)a)->d, y = ((const cand_t*)b)->d;
  return x < y ? -1 : x > y;
}

hnsw_t* hnsw_build(const float* data, int64_t n, int dim, int M, int efc, uint64_t seed) {
  hnsw_t* h = (hnsw_t*)calloc(1, sizeof(hnsw_t));
  h->dim = dim; h->M = M; h->M0 = 2 * M; h->efc = efc; h->n = 0; h->cap = n; h->data = data;
  h->mult = 1.0 / log((double)M); h->rng = seed ? seed : 88172645463325252ull;
  h->level = (int*)calloc((size_t)n, sizeof(int));
  h->links0 = (int*)calloc((size_t)n * (size_t)(h->M0 + 1), sizeof(int));
  h->linksU = (int**)calloc((size_t)n, sizeof(int*));
  h->maxlevel = -1; h->entry = -1;
  unsigned* visited = (unsigned*)calloc((size_t)n, sizeof(unsigned));
  unsigned epoch = 0;
  heap_t res, cands; heap_init(&res, efc + 8); heap_init(&cands, 4 * efc + 64);
  cand_t* sorted = (cand_t*)malloc(sizeof(cand_t) * (size_t)(efc + h->M0 + 8));
  int* sel = (int*)malloc(sizeof(int) * (size_t)(h->M0 + 1));
  for (int64_t id = 0; id < n; ++id) {
    const float* q = data + (size_t)id * dim;
    int lvl = (int)(-log(1.0 - urand(h)) * h->mult);
    h->level[id] = lvl;
    if (lvl > 0) h->linksU[id] = (int*)calloc((size_t)lvl * (size_t)(M + 1), sizeof(int));
    h->n = id + 1;
    if (h->entry < 0) { h->entry = (int)id; h->maxlevel = lvl; continue; }
    int ep = h->entry;
    float epd = l2sqr(q, data + (size_t)ep * dim, dim);
    for (int l = h->maxlevel; l > lvl; --l) {           /* greedy descent, ef = 1 */
      int changed = 1;
      while (changed) {
        changed = 0;
        int* ls = links_of(h, ep, l);
        for (int i = 1; i <= ls[0]; ++i) {
          float d = l2sqr(q, data + (size_t)ls[i] * dim, dim);
          if (d < epd) { epd = d; ep = ls[i]; changed = 1; }
        }
      }
    }
    for (int l = lvl < h->maxlevel ? lvl : h->maxlevel; l >= 0; --l) {
      ++epoch;
      search_layer(h, q, ep, epd, l, efc, &res, &cands, visited, epoch);
      int m = res.n;
      for (int i = 0; i < m; ++i) sorted[i] = res.a[i];
      qsort(sorted, (size_t)m, sizeof(cand_t), cmp_cand);
      ep = sorted[0].id; epd = sorted[0].d;
      const int Mmax = (l == 0) ? h->M0 : M;
      int ns = select_heuristic(h, sorted, m, M, sel);
      int* mine = links_of(h, (int)id, l);
      mine[0] = ns;
      for (int i = 0; i < ns; ++i) mine[1 + i] = sel[i];
      for (int i = 0; i < ns; ++i) {                    /* back links, shrink by heuristic on overflow */
        int nb = sel[i];
        int* ls = links_of(h, nb, l);
        if (ls[0] < Mmax) { ls[++ls[0]] = (int)id; continue; }
        const float* nbv = data + (size_t)nb * dim;
        int cnt = 0;
        sorted[cnt].d = l2sqr(nbv, q, dim); sorted[cnt++].id = (int)id;
        for (int j = 1; j <= ls[0]; ++j) { sorted[cnt].d = l2sqr(nbv, data + (size_t)ls[j] * dim, dim); sorted[cnt++].id = ls[j]; }
        qsort(sorted, (size_t)cnt, sizeof(cand_t), cmp_cand);
        int* tmp = (int*)malloc(sizeof(int) * (size_t)(Mmax + 1));
        int kk = select_heuristic(h, sorted, cnt, Mmax, tmp);
        ls[0] = kk;
        for (int j = 0; j < kk; ++j) ls[1 + j] = tmp[j];
        free(tmp);
      }
    }
    if (lvl > h->maxlevel) { h->maxlevel = lvl; h->entry = (int)id; }
  }
  free(visited); free(res.a); free(cands.a); free(sorted); free(sel);
  return h;
}

/* knn for nq queries; out_ids/out_d are [nq][k] (ascending), -1 padded.  OpenMP over queries. */
void hnsw_search(hnsw_t* h, const float* queries, int64_t nq, int k, int ef, int64_t* out_ids, float* out_d) {
  if (ef < k) ef = k;
#pragma omp parallel
  {
    unsigned* visited = (unsigned*)calloc((size_t)h->n, sizeof(unsigned));
    unsigned epoch = 0;
    heap_t res, cands; heap_init(&res, ef + 8); heap_init(&cands, 4 * ef + 64);
    cand_t* sorted = (cand_t*)malloc(sizeof(cand_t) * (size_t)(ef + 8));
#pragma omp for schedule(dynamic, 4)
    for (int64_t qi = 0; qi < nq; ++qi) {
      const float* q = queries + (size_t)qi * h->dim;
      int ep = h->entry;
      float epd = l2sqr(q, h->data + (size_t)ep * h->dim, h->dim);
      for (int l = h->maxlevel; l > 0; --l) {
        int changed = 1;
        while (changed) {
          changed = 0;
          int* ls = links_of(h, ep, l);
          for (int i = 1; i <= ls[0]; ++i) {
            float d = l2sqr(q, h->data + (size_t)ls[i] * h->dim, h->dim);
            if (d < epd) { epd = d; ep = ls[i]; changed = 1; }
          }
        }
      }
      ++epoch;
      search_layer(h, q, ep, epd, 0, ef, &res, &cands, visited, epoch);
      int m = res.n;
      for (int i = 0; i < m; ++i) sorted[i] = res.a[i];
      qsort(sorted, (size_t)m, sizeof(cand_t), cmp_cand);
      for (int i = 0; i < k; ++i) {
        out_ids[qi * k + i] = i < m ? sorted[i].id : -1;
        out_d[qi * k + i] = i < m ? sorted[i].d : INFINITY;
      }
    }
    free(visited); free(res.a); free(cands.a); free(sorted);
  }
}

void hnsw_free(hnsw_t* h) {
  if (!h) return;
  for (int64_t i = 0; i < h->cap; ++i) free(h->linksU[i]);
  free(h->linksU); free(h->links0); free(h->level); free(h);
}

int hnsw_max_level(const hnsw_t* h) { return h->maxlevel; }
double hnsw_mult(const hnsw_t* h) { return h->mult; }
