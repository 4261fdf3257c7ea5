/*
 * linsys.c -- oracle restatement of TpetraLinearSystem's assembly half:
 * DOF numbering, owned / shared-not-owned CSR graph construction (including the
 * neighbour handshake, simulated in-process for all ranks), sumInto, loadComplete.
 * TEST INFRASTRUCTURE ONLY (see nalu_oracle.h).
 *
 * Follows src/TpetraLinearSystem.C and include/LinearSolver.h:54-120 of the reference.
 * Tpetra itself (CrsGraph storage, Export(ADD)) is un-vendored third-party code; what
 * it does on this path is store the rowptr/colidx Nalu hands it (setAllIndices,
 * TpetraLinearSystem.C:1314-1315) and add shared-not-owned row entries into the owner's
 * row with the same global row/column id (doExport(..., Tpetra::ADD), :1770,1777).
 */
#include "nalu_oracle.h"
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define GID_(gid, ndof, idof) ((int64_t)(ndof) * ((gid)-1) + (idof) + 1) /* TpetraLinearSystem.C:72 */
#define INVALID INT_MAX                                                  /* LinearSolver.h:54 */

typedef struct { int32_t *v; int n, cap; } ivec;
static void ivec_push(ivec *a, int32_t x)
{
  if (a->n == a->cap) {
    a->cap = a->cap ? 2 * a->cap : 8; /* reference reserves 8 (:326) */
    a->v = (int32_t *)realloc(a->v, sizeof(int32_t) * (size_t)a->cap);
  }
  a->v[a->n++] = x;
}
typedef struct { int64_t *v; size_t n, cap; } lvec;
static void lvec_push(lvec *a, int64_t x)
{
  if (a->n == a->cap) {
    a->cap = a->cap ? 2 * a->cap : 64;
    a->v = (int64_t *)realloc(a->v, sizeof(int64_t) * a->cap);
  }
  a->v[a->n++] = x;
}

typedef struct { int64_t gid; int32_t lid; } gl_pair;
static int cmp_gl(const void *a, const void *b)
{
  const gl_pair *x = (const gl_pair *)a, *y = (const gl_pair *)b;
  return (x->gid > y->gid) - (x->gid < y->gid);
}
typedef struct { gl_pair *p; size_t n; } gmap; /* Tpetra::Map::getLocalElement stand-in */
static int32_t gmap_find(const gmap *m, int64_t gid)
{
  size_t lo = 0, hi = m->n;
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if (m->p[mid].gid < gid) lo = mid + 1; else hi = mid;
  }
  return (lo < m->n && m->p[lo].gid == gid) ? m->p[lo].lid : -1;
}

/* LocalGraphArrays (include/LinearSolver.h:62-120) */
typedef struct {
  size_t nrows;
  size_t *rowPointers;
  int32_t *colIndices;
  size_t nnz;
} graph_arrays;

static size_t compute_row_pointers(size_t *rowPtrs, const size_t *rowLens, size_t n)
{
  size_t nnz = 0;
  for (size_t i = 0; i < n; ++i) {
    rowPtrs[i] = nnz;
    nnz += rowLens[i];
  }
  rowPtrs[n] = nnz;
  return nnz;
}
static void graph_init(graph_arrays *g, const size_t *rowLens, size_t n)
{
  g->nrows = n;
  g->rowPointers = (size_t *)malloc(sizeof(size_t) * (n + 1));
  g->nnz = compute_row_pointers(g->rowPointers, rowLens, n);
  g->colIndices = (int32_t *)malloc(sizeof(int32_t) * (g->nnz + 1));
  for (size_t i = 0; i <= g->nnz; ++i) g->colIndices[i] = INVALID;
}
static size_t graph_row_length(const graph_arrays *g, size_t r)
{
  return g->rowPointers[r + 1] - g->rowPointers[r];
}
/* LocalGraphArrays::insertIndices + insert (LinearSolver.h:77-89,107-116) */
static void graph_insert_indices(graph_arrays *g, size_t localRow, size_t numInds,
                                 const int32_t *inds, int numDof)
{
  int32_t *row = g->colIndices + g->rowPointers[localRow];
  size_t rowLen = graph_row_length(g, localRow);
  int32_t *rowEnd = row;
  while (rowEnd < row + rowLen && *rowEnd != INVALID) ++rowEnd; /* std::find(row,row+rowLen,INVALID) */
  for (size_t i = 0; i < numInds; ++i) {
    /* std::lower_bound(row, rowEnd, inds[i]) */
    int32_t *lo = row, *hi = rowEnd;
    while (lo < hi) {
      int32_t *mid = lo + (hi - lo) / 2;
      if (*mid < inds[i]) lo = mid + 1; else hi = mid;
    }
    int32_t *insertPoint = lo;
    if (insertPoint <= rowEnd && *insertPoint != inds[i]) {
      int32_t *newEnd = rowEnd + numDof;
      if (newEnd > row + rowLen) {
        fprintf(stderr, "oracle: graph row overflow (row %zu)\n", localRow);
        abort();
      }
      for (int32_t *ptr = newEnd - 1; ptr != insertPoint; --ptr) {
        if (ptr - numDof >= insertPoint) *ptr = *(ptr - numDof);
      }
      int32_t *ip = insertPoint;
      for (int d = 0; d < numDof; ++d) *ip++ = inds[i] + d;
      rowEnd += numDof;
    }
  }
}

typedef struct {
  /* mesh */
  int32_t nNodes;
  int64_t *gid;
  int32_t *owner;
  int64_t nElem;
  int32_t *elemNodes;
  /* numbering (beginLinearSystemConstruction) */
  int32_t numOwnedNodes, numSharedNotOwned;
  int32_t maxOwnedRowId_, maxSharedNotOwnedRowId_;
  int32_t *ownedAndSharedNodes_; /* representative entity per row node */
  int64_t *ownedGids, *sharedNotOwnedGids; /* dof-expanded row maps */
  int32_t *sharedPids_;
  gmap myLIDs_;      /* entityId -> numDof*localId */
  gmap ownedRowsMap; /* dof gid -> owned row lid */
  gmap sharedRowsMap;
  int32_t *entityToLID_, *entityToColLID_;
  ivec *connections_;
  int inConstruction_;
  /* finalize */
  size_t *ownedRowLengths, *sharedRowLengths;
  lvec ownersAndGids_; /* pairs (owner, gid) flattened */
  int64_t *colGids;    /* totalColsMap_ */
  int32_t nCols;
  gmap colMap;
  graph_arrays ownedGraph, sharedGraph;
  double *ownedVals, *sharedVals, *ownedRhs, *sharedRhs;
} rank_sys;

struct orc_linsys {
  int nRanks;
  int numDof;
  rank_sys *r;
  lvec *sendbuf; /* [src*nRanks+dst] */
};

orc_linsys *orc_linsys_create(int nRanks, int numDof)
{
  orc_linsys *ls = (orc_linsys *)calloc(1, sizeof(orc_linsys));
  ls->nRanks = nRanks;
  ls->numDof = numDof;
  ls->r = (rank_sys *)calloc((size_t)nRanks, sizeof(rank_sys));
  ls->sendbuf = (lvec *)calloc((size_t)nRanks * nRanks, sizeof(lvec));
  return ls;
}

static void graph_free(graph_arrays *g)
{
  free(g->rowPointers);
  free(g->colIndices);
}

void orc_linsys_destroy(orc_linsys *ls)
{
  if (!ls) return;
  for (int p = 0; p < ls->nRanks; ++p) {
    rank_sys *s = &ls->r[p];
    free(s->gid); free(s->owner); free(s->elemNodes);
    free(s->ownedAndSharedNodes_); free(s->ownedGids); free(s->sharedNotOwnedGids);
    free(s->sharedPids_); free(s->myLIDs_.p); free(s->ownedRowsMap.p); free(s->sharedRowsMap.p);
    free(s->entityToLID_); free(s->entityToColLID_);
    if (s->connections_) {
      for (int i = 0; i < s->numOwnedNodes + s->numSharedNotOwned; ++i) free(s->connections_[i].v);
      free(s->connections_);
    }
    free(s->ownedRowLengths); free(s->sharedRowLengths); free(s->ownersAndGids_.v);
    free(s->colGids); free(s->colMap.p);
    graph_free(&s->ownedGraph); graph_free(&s->sharedGraph);
    free(s->ownedVals); free(s->sharedVals); free(s->ownedRhs); free(s->sharedRhs);
  }
  for (int i = 0; i < ls->nRanks * ls->nRanks; ++i) free(ls->sendbuf[i].v);
  free(ls->sendbuf);
  free(ls->r);
  free(ls);
}

void orc_linsys_set_rank_mesh(orc_linsys *ls, int rank, int32_t nNodes, const int64_t *nodeGid,
                              const int32_t *nodeOwner, int64_t nElem, const int32_t *elemNodes)
{
  rank_sys *s = &ls->r[rank];
  s->nNodes = nNodes;
  s->gid = (int64_t *)malloc(sizeof(int64_t) * (size_t)nNodes);
  s->owner = (int32_t *)malloc(sizeof(int32_t) * (size_t)nNodes);
  memcpy(s->gid, nodeGid, sizeof(int64_t) * (size_t)nNodes);
  memcpy(s->owner, nodeOwner, sizeof(int32_t) * (size_t)nNodes);
  s->nElem = nElem;
  s->elemNodes = (int32_t *)malloc(sizeof(int32_t) * 8 * (size_t)nElem);
  memcpy(s->elemNodes, elemNodes, sizeof(int32_t) * 8 * (size_t)nElem);
}

typedef struct { int64_t gid; int32_t node; } gn_pair;
static int cmp_gn(const void *a, const void *b)
{
  const gn_pair *x = (const gn_pair *)a, *y = (const gn_pair *)b;
  if (x->gid != y->gid) return (x->gid > y->gid) - (x->gid < y->gid);
  return (x->node > y->node) - (x->node < y->node);
}

/* TpetraLinearSystem::beginLinearSystemConstruction (:190-327) + fill_entity_to_row_LID_mapping (:954-980) */
static void begin_construction(orc_linsys *ls, int rank)
{
  rank_sys *s = &ls->r[rank];
  if (s->inConstruction_) return;
  s->inConstruction_ = 1;
  const int numDof = ls->numDof;
  /* getDofStatus (:2126-2296) reduces, for a mesh without periodic/overset/non-conformal
   * constraints, to: locally owned -> DS_OwnedDOF, shared and owned elsewhere -> DS_SharedNotOwnedDOF */
  gn_pair *owned = (gn_pair *)malloc(sizeof(gn_pair) * (size_t)(s->nNodes + 1));
  gn_pair *shared = (gn_pair *)malloc(sizeof(gn_pair) * (size_t)(s->nNodes + 1));
  int32_t no = 0, nsh = 0;
  for (int32_t n = 0; n < s->nNodes; ++n) {
    if (s->owner[n] == rank) { owned[no].gid = s->gid[n]; owned[no++].node = n; }
    else { shared[nsh].gid = s->gid[n]; shared[nsh++].node = n; }
  }
  /* std::sort by naluGlobalId then std::unique by id */
  qsort(owned, (size_t)no, sizeof(gn_pair), cmp_gn);
  qsort(shared, (size_t)nsh, sizeof(gn_pair), cmp_gn);
  int32_t nou = 0;
  for (int32_t i = 0; i < no; ++i) if (i == 0 || owned[i].gid != owned[nou - 1].gid) owned[nou++] = owned[i];
  int32_t nsu = 0;
  for (int32_t i = 0; i < nsh; ++i) if (i == 0 || shared[i].gid != shared[nsu - 1].gid) shared[nsu++] = shared[i];
  s->numOwnedNodes = nou;
  s->numSharedNotOwned = nsu;
  s->maxOwnedRowId_ = nou * numDof;
  s->maxSharedNotOwnedRowId_ = (nou + nsu) * numDof;

  s->ownedAndSharedNodes_ = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nou + nsu + 1));
  s->ownedGids = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nou * numDof + 1));
  s->sharedNotOwnedGids = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nsu * numDof + 1));
  s->sharedPids_ = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nsu * numDof + 1));
  s->myLIDs_.p = (gl_pair *)malloc(sizeof(gl_pair) * (size_t)(nou + nsu + 1));
  s->myLIDs_.n = (size_t)(nou + nsu);
  s->ownedRowsMap.p = (gl_pair *)malloc(sizeof(gl_pair) * (size_t)(nou * numDof + 1));
  s->ownedRowsMap.n = (size_t)nou * numDof;
  s->sharedRowsMap.p = (gl_pair *)malloc(sizeof(gl_pair) * (size_t)(nsu * numDof + 1));
  s->sharedRowsMap.n = (size_t)nsu * numDof;

  int32_t localId = 0;
  for (int32_t i = 0; i < nou; ++i) {
    const int64_t entityId = owned[i].gid;
    s->myLIDs_.p[localId].gid = entityId;
    s->myLIDs_.p[localId].lid = numDof * localId;
    s->ownedAndSharedNodes_[localId] = owned[i].node;
    for (int idof = 0; idof < numDof; ++idof) {
      const int64_t gid = GID_(entityId, numDof, idof);
      s->ownedGids[i * numDof + idof] = gid;
      s->ownedRowsMap.p[i * numDof + idof].gid = gid;
      s->ownedRowsMap.p[i * numDof + idof].lid = i * numDof + idof;
    }
    ++localId;
  }
  for (int32_t i = 0; i < nsu; ++i) {
    const int64_t naluId = shared[i].gid;
    s->myLIDs_.p[localId].gid = naluId;
    s->myLIDs_.p[localId].lid = numDof * localId;
    s->ownedAndSharedNodes_[localId] = shared[i].node;
    const int owner = s->owner[shared[i].node];
    for (int idof = 0; idof < numDof; ++idof) {
      const int64_t gid = GID_(naluId, numDof, idof);
      s->sharedNotOwnedGids[i * numDof + idof] = gid;
      s->sharedPids_[i * numDof + idof] = owner;
      s->sharedRowsMap.p[i * numDof + idof].gid = gid;
      s->sharedRowsMap.p[i * numDof + idof].lid = i * numDof + idof;
    }
    ++localId;
  }
  free(owned);
  free(shared);
  /* both halves of myLIDs_ are gid-sorted, but owned and shared gids interleave: sort the union */
  qsort(s->myLIDs_.p, s->myLIDs_.n, sizeof(gl_pair), cmp_gl);
  /* row maps are gid-sorted by construction */

  s->entityToLID_ = (int32_t *)malloc(sizeof(int32_t) * (size_t)(s->nNodes + 1));
  for (int32_t n = 0; n < s->nNodes; ++n) {
    int32_t lid = gmap_find(&s->myLIDs_, s->gid[n]);
    s->entityToLID_[n] = lid < 0 ? 2000000000 : lid;
  }
  s->connections_ = (ivec *)calloc((size_t)(nou + nsu + 1), sizeof(ivec));
}

/* TpetraLinearSystem::insert_connection (:329-349) */
static void insert_connection(orc_linsys *ls, rank_sys *s, int32_t a, int32_t b)
{
  size_t idx = (size_t)(s->entityToLID_[a] / ls->numDof);
  ivec *vec = &s->connections_[idx];
  for (int i = 0; i < vec->n; ++i) if (vec->v[i] == b) return; /* std::find */
  ivec_push(vec, b);
}

/* TpetraLinearSystem::addConnections (:351-367) */
static void add_connections(orc_linsys *ls, rank_sys *s, const int32_t *entities, size_t num_entities)
{
  for (size_t a = 0; a < num_entities; ++a) {
    const int32_t entity_a = entities[a];
    const int64_t id_a = s->gid[entity_a];
    insert_connection(ls, s, entity_a, entity_a);
    for (size_t b = a + 1; b < num_entities; ++b) {
      const int32_t entity_b = entities[b];
      const int64_t id_b = s->gid[entity_b];
      const int a_then_b = id_a < id_b;
      const int32_t entity_min = a_then_b ? entity_a : entity_b;
      const int32_t entity_max = a_then_b ? entity_b : entity_a;
      insert_connection(ls, s, entity_min, entity_max);
    }
  }
}

static int cmp_i32(const void *a, const void *b)
{
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}
typedef struct { int64_t owner, gid; } og_pair;
static int cmp_og(const void *a, const void *b)
{
  const og_pair *x = (const og_pair *)a, *y = (const og_pair *)b;
  if (x->owner != y->owner) return (x->owner > y->owner) - (x->owner < y->owner);
  return (x->gid > y->gid) - (x->gid < y->gid);
}

/* add_to_length (:666-677) */
static void add_to_length(rank_sys *s, unsigned numDof, int32_t lid_a, int a_owned, unsigned numColEntities)
{
  size_t *v = a_owned ? s->ownedRowLengths : s->sharedRowLengths;
  int32_t lid = a_owned ? lid_a : lid_a - s->maxOwnedRowId_;
  for (unsigned d = 0; d < numDof; ++d) v[lid + d] += (size_t)numDof * numColEntities;
}

/* add_lengths_to_comm (:679-698) */
static void add_lengths_to_comm(orc_linsys *ls, int src, int owner, int64_t entityId_a, unsigned numDof,
                                unsigned numColEntities, const int64_t *colEntityIds, const int *colOwners)
{
  lvec *sbuf = &ls->sendbuf[src * ls->nRanks + owner];
  lvec_push(sbuf, GID_(entityId_a, numDof, 0));
  lvec_push(sbuf, (int64_t)numColEntities * 2);
  for (unsigned c = 0; c < numColEntities; ++c) {
    lvec_push(sbuf, GID_(colEntityIds[c], numDof, 0));
    lvec_push(sbuf, colOwners[c]);
  }
}

void orc_linsys_build_elem_to_node_graph(orc_linsys *ls)
{
  /* buildElemToNodeGraph (:415-420) -> buildConnectedNodeGraph (:392-413) */
  for (int p = 0; p < ls->nRanks; ++p) {
    rank_sys *s = &ls->r[p];
    begin_construction(ls, p);
    for (int64_t e = 0; e < s->nElem; ++e) add_connections(ls, s, s->elemNodes + 8 * e, 8);
  }
}

/* TpetraLinearSystem::finalizeLinearSystem (:1257-1353) */
void orc_linsys_finalize(orc_linsys *ls)
{
  const int numDof = ls->numDof;
  const int P = ls->nRanks;
  /* phase 1 on every rank: sort, row lengths, pack messages, storeOwnersForShared */
  for (int p = 0; p < P; ++p) {
    rank_sys *s = &ls->r[p];
    begin_construction(ls, p);
    s->inConstruction_ = 0;
    const int nRowNodes = s->numOwnedNodes + s->numSharedNotOwned;
    /* sort_connections (:659-664): by entity handle */
    for (int i = 0; i < nRowNodes; ++i)
      qsort(s->connections_[i].v, (size_t)s->connections_[i].n, sizeof(int32_t), cmp_i32);
    s->ownedRowLengths = (size_t *)calloc((size_t)(s->maxOwnedRowId_ + 1), sizeof(size_t));
    s->sharedRowLengths = (size_t *)calloc((size_t)(s->numSharedNotOwned * numDof + 1), sizeof(size_t));

    /* compute_graph_row_lengths (:803-866) */
    int64_t *colEntityIds = (int64_t *)malloc(sizeof(int64_t) * 4096);
    int *colOwners = (int *)malloc(sizeof(int) * 4096);
    for (int i = 0; i < nRowNodes; ++i) {
      const ivec *colEntities = &s->connections_[i];
      const unsigned numColEntities = (unsigned)colEntities->n;
      const int32_t entity_a = s->ownedAndSharedNodes_[i];
      for (unsigned j = 0; j < numColEntities; ++j) {
        colEntityIds[j] = s->gid[colEntities->v[j]];
        colOwners[j] = s->owner[colEntities->v[j]];
      }
      const int64_t entityId_a = s->gid[entity_a];
      const int entity_a_owned = s->owner[entity_a] == p;
      const int32_t lid_a = s->entityToLID_[entity_a];
      int entity_a_owner = s->owner[entity_a];

      add_to_length(s, (unsigned)numDof, lid_a, entity_a_owned, numColEntities);

      const int entity_a_shared = !entity_a_owned;
      if (entity_a_shared)
        add_lengths_to_comm(ls, p, entity_a_owner, entityId_a, (unsigned)numDof, numColEntities, colEntityIds, colOwners);

      for (unsigned ii = 0; ii < numColEntities; ++ii) {
        const int32_t entity_b = colEntities->v[ii];
        if (entity_b == entity_a) continue;
        const int64_t entityId_b = colEntityIds[ii];
        const int entity_b_owned = s->owner[entity_b] == p;
        const int32_t lid_b = s->entityToLID_[entity_b];
        add_to_length(s, (unsigned)numDof, lid_b, entity_b_owned, 1);
        if (!entity_b_owned)
          add_lengths_to_comm(ls, p, colOwners[ii], entityId_b, (unsigned)numDof, 1, &entityId_a, &entity_a_owner);
      }
    }
    free(colEntityIds);
    free(colOwners);

    /* storeOwnersForShared (:1001-1021) */
    s->ownersAndGids_.n = 0;
    for (int32_t n = 0; n < s->nNodes; ++n) {
      if (s->owner[n] != p) {
        for (int idof = 0; idof < numDof; ++idof) {
          lvec_push(&s->ownersAndGids_, s->owner[n]);
          lvec_push(&s->ownersAndGids_, GID_(s->gid[n], numDof, idof));
        }
      }
    }
  }

  /* communicate_remote_columns (:700-738): receiver side, neighbours in ascending rank */
  for (int p = 0; p < P; ++p) {
    rank_sys *s = &ls->r[p];
    for (int q = 0; q < P; ++q) {
      const lvec *rbuf = &ls->sendbuf[q * P + p];
      size_t pos = 0;
      while (pos < rbuf->n) {
        const int64_t rowGid = rbuf->v[pos++];
        const unsigned len = (unsigned)rbuf->v[pos++];
        const unsigned numCols = len / 2;
        int32_t lid = gmap_find(&s->ownedRowsMap, rowGid);
        if (lid < 0) { fprintf(stderr, "oracle: P%d lid<0 for rowGid=%lld from %d\n", p, (long long)rowGid, q); abort(); }
        for (int d = 0; d < numDof; ++d) s->ownedRowLengths[lid++] += (size_t)numCols * numDof;
        for (unsigned i = 0; i < numCols; ++i) {
          int64_t colGid = rbuf->v[pos++];
          const int64_t owner = rbuf->v[pos++];
          for (int dd = 0; dd < numDof; ++dd) {
            lvec_push(&s->ownersAndGids_, owner);
            lvec_push(&s->ownersAndGids_, colGid++);
          }
        }
      }
    }
  }

  for (int p = 0; p < P; ++p) {
    rank_sys *s = &ls->r[p];
    const int nRowNodes = s->numOwnedNodes + s->numSharedNotOwned;
    graph_init(&s->ownedGraph, s->ownedRowLengths, (size_t)s->maxOwnedRowId_);
    graph_init(&s->sharedGraph, s->sharedRowLengths, (size_t)s->numSharedNotOwned * numDof);

    /* fill_owned_and_shared_then_nonowned_ordered_by_proc (:1051-1094); ownersAndGids_ is a std::set */
    size_t nog = s->ownersAndGids_.n / 2;
    og_pair *og = (og_pair *)malloc(sizeof(og_pair) * (nog + 1));
    for (size_t i = 0; i < nog; ++i) { og[i].owner = s->ownersAndGids_.v[2 * i]; og[i].gid = s->ownersAndGids_.v[2 * i + 1]; }
    qsort(og, nog, sizeof(og_pair), cmp_og);
    size_t nu = 0;
    for (size_t i = 0; i < nog; ++i)
      if (i == 0 || og[i].owner != og[nu - 1].owner || og[i].gid != og[nu - 1].gid) og[nu++] = og[i];
    size_t nOwnedG = (size_t)s->maxOwnedRowId_, nSharedG = (size_t)s->numSharedNotOwned * numDof;
    s->colGids = (int64_t *)malloc(sizeof(int64_t) * (nOwnedG + nSharedG + nu + 1));
    size_t nc = 0;
    for (size_t i = 0; i < nOwnedG; ++i) s->colGids[nc++] = s->ownedGids[i];
    for (size_t i = 0; i < nSharedG; ++i) s->colGids[nc++] = s->sharedNotOwnedGids[i];
    for (size_t i = 0; i < nu; ++i) {
      if (og[i].owner != p && gmap_find(&s->ownedRowsMap, og[i].gid) < 0 &&
          gmap_find(&s->sharedRowsMap, og[i].gid) < 0)
        s->colGids[nc++] = og[i].gid;
    }
    free(og);
    s->nCols = (int32_t)nc;
    s->colMap.p = (gl_pair *)malloc(sizeof(gl_pair) * (nc + 1));
    s->colMap.n = nc;
    for (size_t i = 0; i < nc; ++i) { s->colMap.p[i].gid = s->colGids[i]; s->colMap.p[i].lid = (int32_t)i; }
    qsort(s->colMap.p, nc, sizeof(gl_pair), cmp_gl);

    /* fill_entity_to_col_LID_mapping (:982-999) */
    s->entityToColLID_ = (int32_t *)malloc(sizeof(int32_t) * (size_t)(s->nNodes + 1));
    for (int32_t n = 0; n < s->nNodes; ++n)
      s->entityToColLID_[n] = gmap_find(&s->colMap, GID_(s->gid[n], numDof, 0));

    /* insert_graph_connections (:877-922) */
    int32_t *localDofs_b = (int32_t *)malloc(sizeof(int32_t) * 4096);
    for (int i = 0; i < nRowNodes; ++i) {
      const ivec *entities_b = &s->connections_[i];
      const unsigned numColEntities = (unsigned)entities_b->n;
      if (numColEntities > 0) {
        const int32_t entity_a = s->ownedAndSharedNodes_[i];
        const int a_owned = s->owner[entity_a] == p;
        int32_t localDofs_a = s->entityToColLID_[entity_a];
        for (unsigned j = 0; j < numColEntities; ++j) localDofs_b[j] = s->entityToColLID_[entities_b->v[j]];
        {
          graph_arrays *g = a_owned ? &s->ownedGraph : &s->sharedGraph;
          int32_t rowLid = s->entityToLID_[entity_a];
          if (rowLid >= s->maxOwnedRowId_) rowLid -= s->maxOwnedRowId_;
          graph_insert_indices(g, (size_t)rowLid, numColEntities, localDofs_b, numDof);
        }
        for (unsigned j = 0; j < numColEntities; ++j) {
          if (entities_b->v[j] != entity_a) {
            const int b_owned = s->owner[entities_b->v[j]] == p;
            graph_arrays *g = b_owned ? &s->ownedGraph : &s->sharedGraph;
            int32_t rowLid = s->entityToLID_[entities_b->v[j]];
            if (rowLid >= s->maxOwnedRowId_) rowLid -= s->maxOwnedRowId_;
            graph_insert_indices(g, (size_t)rowLid, 1, &localDofs_a, numDof);
          }
        }
      }
    }
    free(localDofs_b);

    /* insert_communicated_col_indices (:924-952) */
    int32_t *colLids = (int32_t *)malloc(sizeof(int32_t) * 4096);
    for (int q = 0; q < P; ++q) {
      const lvec *rbuf = &ls->sendbuf[q * P + p];
      size_t pos = 0;
      while (pos < rbuf->n) {
        const int64_t rowGid = rbuf->v[pos++];
        const unsigned len = (unsigned)rbuf->v[pos++];
        const unsigned numCols = len / 2;
        int32_t rowLid = gmap_find(&s->ownedRowsMap, rowGid);
        for (unsigned i = 0; i < numCols; ++i) {
          const int64_t colGid = rbuf->v[pos++];
          pos++; /* owner */
          colLids[i] = gmap_find(&s->colMap, colGid);
        }
        graph_insert_indices(&s->ownedGraph, (size_t)rowLid, numCols, colLids, numDof);
      }
    }
    free(colLids);

    /* fill_in_extra_dof_rows_per_node (:1227-1246) */
    for (int which = 0; which < 2; ++which) {
      graph_arrays *g = which ? &s->sharedGraph : &s->ownedGraph;
      if (numDof == 1) continue;
      for (size_t i = 0; i < g->nrows;) {
        const int32_t *row = g->colIndices + g->rowPointers[i];
        size_t rowLen = graph_row_length(g, i);
        for (int d = 1; d < numDof; ++d) {
          int32_t *row_d = g->colIndices + g->rowPointers[i] + rowLen * (size_t)d;
          for (size_t j = 0; j < rowLen; ++j) row_d[j] = row[j];
        }
        i += (size_t)numDof;
      }
    }

    /* remove_invalid_indices (:1193-1225): owned graph only */
    {
      graph_arrays *g = &s->ownedGraph;
      size_t nnz = g->rowPointers[g->nrows];
      size_t newNnz = 0;
      for (size_t i = 0; i < g->nrows; ++i) {
        const int32_t *row = g->colIndices + g->rowPointers[i];
        int rowLen = (int)graph_row_length(g, i);
        for (int j = rowLen - 1; j >= 0; --j) {
          if (row[j] != INVALID) { s->ownedRowLengths[i] = (size_t)j + 1; break; }
        }
        newNnz += s->ownedRowLengths[i];
      }
      if (newNnz < nnz) {
        int32_t *newCols = (int32_t *)malloc(sizeof(int32_t) * (newNnz + 1));
        size_t index = 0;
        for (size_t i = 0; i < g->nrows; ++i) {
          const int32_t *row = g->colIndices + g->rowPointers[i];
          for (size_t j = 0; j < s->ownedRowLengths[i]; ++j) newCols[index++] = row[j];
        }
        free(g->colIndices);
        g->colIndices = newCols;
        g->nnz = compute_row_pointers(g->rowPointers, s->ownedRowLengths, g->nrows);
      }
    }

    s->ownedVals = (double *)calloc(s->ownedGraph.nnz + 1, sizeof(double));
    s->sharedVals = (double *)calloc(s->sharedGraph.nnz + 1, sizeof(double));
    s->ownedRhs = (double *)calloc((size_t)s->maxOwnedRowId_ + 1, sizeof(double));
    s->sharedRhs = (double *)calloc((size_t)s->numSharedNotOwned * numDof + 1, sizeof(double));
  }
}

int32_t orc_linsys_num_owned_rows(const orc_linsys *ls, int rank) { return ls->r[rank].maxOwnedRowId_; }
int32_t orc_linsys_num_shared_rows(const orc_linsys *ls, int rank) { return ls->r[rank].numSharedNotOwned * ls->numDof; }
int32_t orc_linsys_num_cols(const orc_linsys *ls, int rank) { return ls->r[rank].nCols; }
const size_t *orc_linsys_rowptr(const orc_linsys *ls, int rank, int shared)
{ return shared ? ls->r[rank].sharedGraph.rowPointers : ls->r[rank].ownedGraph.rowPointers; }
const int32_t *orc_linsys_colidx(const orc_linsys *ls, int rank, int shared)
{ return shared ? ls->r[rank].sharedGraph.colIndices : ls->r[rank].ownedGraph.colIndices; }
double *orc_linsys_values(orc_linsys *ls, int rank, int shared)
{ return shared ? ls->r[rank].sharedVals : ls->r[rank].ownedVals; }
double *orc_linsys_rhs(orc_linsys *ls, int rank, int shared)
{ return shared ? ls->r[rank].sharedRhs : ls->r[rank].ownedRhs; }
const int64_t *orc_linsys_row_gids(const orc_linsys *ls, int rank, int shared)
{ return shared ? ls->r[rank].sharedNotOwnedGids : ls->r[rank].ownedGids; }
const int64_t *orc_linsys_col_gids(const orc_linsys *ls, int rank) { return ls->r[rank].colGids; }
const int32_t *orc_linsys_node_row_lid(const orc_linsys *ls, int rank) { return ls->r[rank].entityToLID_; }
const int32_t *orc_linsys_node_col_lid(const orc_linsys *ls, int rank) { return ls->r[rank].entityToColLID_; }

/* TpetraLinearSystem::zeroSystem (:1355-1372) */
void orc_linsys_zero(orc_linsys *ls)
{
  for (int p = 0; p < ls->nRanks; ++p) {
    rank_sys *s = &ls->r[p];
    memset(s->ownedVals, 0, sizeof(double) * s->ownedGraph.nnz);
    memset(s->sharedVals, 0, sizeof(double) * s->sharedGraph.nnz);
    memset(s->ownedRhs, 0, sizeof(double) * (size_t)s->maxOwnedRowId_);
    memset(s->sharedRhs, 0, sizeof(double) * (size_t)s->numSharedNotOwned * ls->numDof);
  }
}

/* Tpetra::Details::shellSortKeysAndValues (Tpetra_Util.hpp; un-vendored): Shell sort, gap halves */
void orc_shell_sort_keys_and_values(int *keys, int *values, int n)
{
  int midpoint = n / 2;
  while (midpoint > 0) {
    const int max = n - midpoint;
    for (int j = 0; j < max; ++j) {
      for (int k = j; k >= 0; k -= midpoint) {
        if (keys[k + midpoint] >= keys[k]) break;
        int tk = keys[k + midpoint]; keys[k + midpoint] = keys[k]; keys[k] = tk;
        int tv = values[k + midpoint]; values[k + midpoint] = values[k]; values[k] = tv;
      }
    }
    midpoint = midpoint / 2;
  }
}

/* sum_into_row_vec_3 (:1376-1413) */
static void sum_into_row_vec_3(const int32_t *colidx, double *value, int length, int num_entities,
                               const int *localIds, const int *sort_permutation, const double *input_values)
{
  int offset = 0;
  for (int j = 0; j < num_entities; ++j) {
    const int id_index = 3 * j;
    const int cur_local_column_idx = localIds[id_index];
    while (colidx[offset] != cur_local_column_idx) {
      offset += 3;
      if (offset >= length) return;
    }
    const int entry_offset = sort_permutation[id_index];
    value[offset + 0] += input_values[entry_offset + 0];
    value[offset + 1] += input_values[entry_offset + 1];
    value[offset + 2] += input_values[entry_offset + 2];
    offset += 3;
  }
}

/* sum_into_row (:1415-1453) */
static void sum_into_row(const int32_t *colidx, double *value, int length, int num_entities, int numDof,
                         const int *localIds, const int *sort_permutation, const double *input_values)
{
  if (numDof == 3) {
    sum_into_row_vec_3(colidx, value, length, num_entities, localIds, sort_permutation, input_values);
    return;
  }
  const int numCols = num_entities * numDof;
  int offset = 0;
  for (int j = 0; j < numCols; ++j) {
    const int perm_index = sort_permutation[j];
    const int cur_local_column_idx = localIds[j];
    while (offset < length && colidx[offset] != cur_local_column_idx) ++offset;
    if (offset < length) value[offset] += input_values[perm_index];
  }
}

/* TpetraLinearSystem::sumInto (:1457-1522) */
void orc_linsys_sum_into(orc_linsys *ls, int rank, int numEntities, const int32_t *entities,
                         const double *rhs, const double *lhs)
{
  rank_sys *s = &ls->r[rank];
  const int numDof_ = ls->numDof;
  const int n_obj = numEntities;
  const int numRows = n_obj * numDof_;
  int localIds[81 * 3], sortPermutation[81 * 3];

  for (int i = 0; i < n_obj; i++) {
    const int localOffset = s->entityToColLID_[entities[i]];
    for (int d = 0; d < numDof_; ++d) localIds[i * numDof_ + d] = localOffset + d;
  }
  for (int i = 0; i < numRows; ++i) sortPermutation[i] = i;
  orc_shell_sort_keys_and_values(localIds, sortPermutation, numRows);

  for (int r = 0; r < numRows; ++r) {
    int i = sortPermutation[r] / numDof_;
    int rowLid = s->entityToLID_[entities[i]];
    rowLid += sortPermutation[r] % numDof_;
    const int cur_perm_index = sortPermutation[r];
    const double *cur_lhs = lhs + (size_t)cur_perm_index * numRows;
    const double cur_rhs = rhs[cur_perm_index];

    if (rowLid < s->maxOwnedRowId_) {
      const size_t b = s->ownedGraph.rowPointers[rowLid];
      const int len = (int)(s->ownedGraph.rowPointers[rowLid + 1] - b);
      sum_into_row(s->ownedGraph.colIndices + b, s->ownedVals + b, len, n_obj, numDof_, localIds, sortPermutation, cur_lhs);
      s->ownedRhs[rowLid] += cur_rhs;
    } else if (rowLid < s->maxSharedNotOwnedRowId_) {
      const int actualLocalId = rowLid - s->maxOwnedRowId_;
      const size_t b = s->sharedGraph.rowPointers[actualLocalId];
      const int len = (int)(s->sharedGraph.rowPointers[actualLocalId + 1] - b);
      sum_into_row(s->sharedGraph.colIndices + b, s->sharedVals + b, len, n_obj, numDof_, localIds, sortPermutation, cur_lhs);
      s->sharedRhs[actualLocalId] += cur_rhs;
    }
  }
}

/* AssembleElemSolverAlgorithm::execute (src/AssembleElemSolverAlgorithm.C:74-100):
 * every locally-owned element: gather + ME + kernels, then apply_coeff -> sumInto */
void orc_assemble_elem(orc_linsys *ls, int rank, const orc_kernel_desc *kernels, int nk)
{
  rank_sys *s = &ls->r[rank];
  const int numDof = ls->numDof;
  const int rhsSize = 8 * numDof;
  double *lhs = (double *)malloc(sizeof(double) * (size_t)rhsSize * rhsSize);
  double *rhs = (double *)malloc(sizeof(double) * (size_t)rhsSize);
  for (int64_t e = 0; e < s->nElem; ++e) {
    const int32_t *nodes = s->elemNodes + 8 * e;
    orc_element_lhs_rhs(kernels, nk, numDof, nodes, e, lhs, rhs);
    orc_linsys_sum_into(ls, rank, 8, nodes, rhs, lhs);
  }
  free(lhs);
  free(rhs);
}

/* AssembleNodeSolverAlgorithm::execute (src/AssembleNodeSolverAlgorithm.C:57-111) with the
 * {Momentum,Continuity,Scalar}MassBDF2NodeSuppAlg::node_execute bodies
 * (src/MomentumMassBDF2NodeSuppAlg.C:80-103, src/ContinuityMassBDF2NodeSuppAlg.C:73-89,
 *  src/ScalarMassBDF2NodeSuppAlg.C:82-100); locally owned nodes only */
void orc_assemble_node_mass(orc_linsys *ls, int rank, const orc_kernel_desc *k, const double *dualNodalVolume)
{
  rank_sys *s = &ls->r[rank];
  const int numDof = ls->numDof;
  const double gamma1_ = k->gamma1, gamma2_ = k->gamma2, gamma3_ = k->gamma3, dt_ = k->dt;
  double lhs[9], rhs[3];
  for (int32_t node = 0; node < s->nNodes; ++node) {
    if (s->owner[node] != rank) continue;
    for (int i = 0; i < numDof * numDof; ++i) lhs[i] = 0.0;
    for (int i = 0; i < numDof; ++i) rhs[i] = 0.0;
    const double dualVolume = dualNodalVolume[node];
    if (k->kind == ORC_MOMENTUM_MASS) {
      const double *uNm1 = k->velocityNm1 + 3 * (size_t)node;
      const double *uN = k->velocityN + 3 * (size_t)node;
      const double *uNp1 = k->velocity + 3 * (size_t)node;
      const double rhoNm1 = k->densityNm1[node], rhoN = k->densityN[node], rhoNp1 = k->density[node];
      const double *dpdx = k->Gpdx + 3 * (size_t)node;
      const double lhsfac = gamma1_ * rhoNp1 * dualVolume / dt_;
      for (int i = 0; i < 3; ++i) {
        rhs[i] += -(gamma1_ * rhoNp1 * uNp1[i] + gamma2_ * rhoN * uN[i] + gamma3_ * rhoNm1 * uNm1[i]) * dualVolume / dt_ -
                  dpdx[i] * dualVolume;
        const int row = i * 3;
        lhs[row + i] += lhsfac;
      }
    } else if (k->kind == ORC_CONTINUITY_MASS) {
      const double projTimeScale = dt_ / gamma1_;
      const double rhoNm1 = k->densityNm1[node], rhoN = k->densityN[node], rhoNp1 = k->density[node];
      rhs[0] -= (gamma1_ * rhoNp1 + gamma2_ * rhoN + gamma3_ * rhoNm1) * dualVolume / dt_ / projTimeScale;
      lhs[0] += 0.0;
    } else if (k->kind == ORC_SCALAR_MASS) {
      const double qNm1 = k->scalarQNm1[node], qN = k->scalarQN[node], qNp1 = k->scalarQ[node];
      const double rhoNm1 = k->densityNm1[node], rhoN = k->densityN[node], rhoNp1 = k->density[node];
      const double lhsTime = gamma1_ * rhoNp1 * dualVolume / dt_;
      rhs[0] -= (gamma1_ * rhoNp1 * qNp1 + gamma2_ * qN * rhoN + gamma3_ * qNm1 * rhoNm1) * dualVolume / dt_;
      lhs[0] += lhsTime;
    }
    orc_linsys_sum_into(ls, rank, 1, &node, rhs, lhs);
  }
}

/* TpetraLinearSystem::loadComplete (:1757-1778): doExport(sharedNotOwned -> owned, Tpetra::ADD) */
void orc_linsys_load_complete(orc_linsys *ls)
{
  const int numDof = ls->numDof;
  for (int p = 0; p < ls->nRanks; ++p) {
    rank_sys *s = &ls->r[p];
    const int nShared = s->numSharedNotOwned * numDof;
    for (int r = 0; r < nShared; ++r) {
      const int64_t rowGid = s->sharedNotOwnedGids[r];
      const int owner = s->sharedPids_[r];
      rank_sys *t = &ls->r[owner];
      const int32_t tl = gmap_find(&t->ownedRowsMap, rowGid);
      if (tl < 0) { fprintf(stderr, "oracle: export target row missing\n"); abort(); }
      const size_t tb = t->ownedGraph.rowPointers[tl], te = t->ownedGraph.rowPointers[tl + 1];
      for (size_t j = s->sharedGraph.rowPointers[r]; j < s->sharedGraph.rowPointers[r + 1]; ++j) {
        const int32_t c = s->sharedGraph.colIndices[j];
        if (c == INVALID) continue;
        const int64_t colGid = s->colGids[c];
        const int32_t tc = gmap_find(&t->colMap, colGid);
        /* owner row is sorted by column lid */
        size_t lo = tb, hi = te;
        while (lo < hi) {
          size_t mid = (lo + hi) / 2;
          if (t->ownedGraph.colIndices[mid] < tc) lo = mid + 1; else hi = mid;
        }
        if (lo < te && t->ownedGraph.colIndices[lo] == tc) t->ownedVals[lo] += s->sharedVals[j];
        else { fprintf(stderr, "oracle: export target column missing\n"); abort(); }
      }
      t->ownedRhs[tl] += s->sharedRhs[r];
    }
  }
}

/* thin test hooks for LocalGraphArrays (unit_tests/UnitTestLocalGraphArrays.C) */
void *orc_graph_arrays_create(const size_t *rowLens, size_t n)
{
  graph_arrays *g = (graph_arrays *)calloc(1, sizeof(graph_arrays));
  graph_init(g, rowLens, n);
  return g;
}
void orc_graph_arrays_insert(void *h, size_t localRow, size_t numInds, const int32_t *inds, int numDof)
{
  graph_insert_indices((graph_arrays *)h, localRow, numInds, inds, numDof);
}
const int32_t *orc_graph_arrays_colidx(void *h) { return ((graph_arrays *)h)->colIndices; }
const size_t *orc_graph_arrays_rowptr(void *h) { return ((graph_arrays *)h)->rowPointers; }
void orc_graph_arrays_destroy(void *h)
{
  graph_free((graph_arrays *)h);
  free(h);
}
