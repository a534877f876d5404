/*
 * pr_oracle_poly.c — CPU oracle for the reference's postProcessPlanes re-absorption pass.
 * TEST INFRASTRUCTURE ONLY (see pr_oracle.h).  Restates Dialog/PlaneDetect.h:1442-1448 (projPoint2Plane),
 * :203-207 (distP2P), :2019-2023 (getInfoBetPointAndPlane), :1891-1955 (isPointInPoly), :1957-2016
 * (isBothLineSegsIntersect) and the claim / peel loop :1530-1566.  Pinned against the reference's own source for
 * these functions compiled over a minimal type shim (oracle/build_ref.py -> oracle/_ref), see tests/test_reabsorb.py.
 */
#include "pr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float v[3]; } vec3;

/* Eigen: Vector3f::dot / squaredNorm -> redux over 3 coefficients = e0 + (e1 + e2) (CHOICE, see header) */
static inline float dot3(const vec3* a, const vec3* b) {
  return a->v[0] * b->v[0] + (a->v[1] * b->v[1] + a->v[2] * b->v[2]);
}

/* Eigen >= 3.3 normalize(): z = squaredNorm(); if (z > 0) *this /= sqrt(z) */
static inline void normalize3(vec3* a) {
  const float z = dot3(a, a);
  if (z > 0.0f) {
    const float n = sqrtf(z);
    a->v[0] /= n;
    a->v[1] /= n;
    a->v[2] /= n;
  }
}

static inline vec3 cross3(const vec3* a, const vec3* b) {
  vec3 r;
  r.v[0] = a->v[1] * b->v[2] - a->v[2] * b->v[1];
  r.v[1] = a->v[2] * b->v[0] - a->v[0] * b->v[2];
  r.v[2] = a->v[0] * b->v[1] - a->v[1] * b->v[0];
  return r;
}

/* distP2P (:203-207): a = dx*dx + dy*dy + dz*dz left to right, pow(a, 0.5f) -> sqrtf (CHOICE) */
static inline float dist_p2p(const orc_point* p1, const orc_point* p2) {
  const float a = (p1->x - p2->x) * (p1->x - p2->x) + (p1->y - p2->y) * (p1->y - p2->y) + (p1->z - p2->z) * (p1->z - p2->z);
  return sqrtf(a);
}

/* projPoint2Plane (:1442-1448): lambda is a float; lambda / 2.0 * n and the subtraction run in double */
static inline void proj_point(const orc_point* s, const float c[4], orc_point* d) {
  float lambda = 2.0 * (c[0] * s->x + c[1] * s->y + c[2] * s->z + c[3]);
  d->x = s->x - lambda / 2.0 * c[0];
  d->y = s->y - lambda / 2.0 * c[1];
  d->z = s->z - lambda / 2.0 * c[2];
  d->w = 1.0f;
}

void orc_msvc_rand_edges(unsigned seed, int border_size, int edges[10]) {
  unsigned long hold = seed; /* srand(seed) */
  for (int i = 0; i < 10; ++i) {
    hold = (hold * 214013ul + 2531011ul) & 0xFFFFFFFFul;
    const int r = (int)((hold >> 16) & 0x7fff);
    edges[i] = (int)((unsigned)r % (unsigned)border_size);
  }
}

/* isBothLineSegsIntersect (:1957-2016) */
int orc_segs_intersect(const orc_point* pa, const orc_point* pb, const orc_point* pc, const orc_point* pd) {
  vec3 nab = {{pb->x - pa->x, pb->y - pa->y, pb->z - pa->z}};
  normalize3(&nab);
  vec3 ncd = {{pd->x - pc->x, pd->y - pc->y, pd->z - pc->z}};
  normalize3(&ncd);
  const vec3 pa_pc = {{pc->x - pa->x, pc->y - pa->y, pc->z - pa->z}};
  float lambda1, lambda2;
  const float nn = dot3(&nab, &ncd);
  if (fabsf(nn) <= 0.001f) {
    lambda1 = dot3(&nab, &pa_pc);
    lambda2 = -1.0f * dot3(&ncd, &pa_pc);
  } else if (nn >= 0.9999f) {
    return 0;
  } else {
    const float c1 = 1.0f - nn * nn;
    const float c2 = dot3(&nab, &pa_pc) * nn - dot3(&ncd, &pa_pc);
    lambda2 = c2 / c1;
    lambda1 = (lambda2 + dot3(&ncd, &pa_pc)) / nn;
  }
  orc_point p1, p2, pi;
  p1.x = pa->x + lambda1 * nab.v[0];
  p1.y = pa->y + lambda1 * nab.v[1];
  p1.z = pa->z + lambda1 * nab.v[2];
  p2.x = pc->x + lambda2 * ncd.v[0];
  p2.y = pc->y + lambda2 * ncd.v[1];
  p2.z = pc->z + lambda2 * ncd.v[2];
  pi.x = (p1.x + p2.x) / 2.0f;
  pi.y = (p1.y + p2.y) / 2.0f;
  pi.z = (p1.z + p2.z) / 2.0f;
  const float dist_pa = dist_p2p(&pi, pa), dist_pb = dist_p2p(&pi, pb), dist_ab = dist_p2p(pa, pb);
  const float dist_pc = dist_p2p(&pi, pc), dist_pd = dist_p2p(&pi, pd), dist_cd = dist_p2p(pc, pd);
  return fabsf(dist_pa + dist_pb - dist_ab) < 0.001f && fabsf(dist_pc + dist_pd - dist_cd) < 0.001f;
}

/* isPointInPoly (:1891-1955) */
int orc_point_in_poly(const orc_point* p, const float coeff[4], const orc_point* border, int nb, float t, unsigned seed) {
  orc_point p_proj;
  proj_point(p, coeff, &p_proj);
  const float dist = dist_p2p(p, &p_proj);
  if (dist > t) return 0;
  int edges[10];
  orc_msvc_rand_edges(seed, nb, edges);
  const float lambda = 10000;
  const vec3 plane_norm = {{coeff[0], coeff[1], coeff[2]}};
  int odd = 0;
  for (int i = 0; i < 10; ++i) {
    const int index = edges[i];
    const orc_point* sp = &border[index];
    const orc_point* ep = &border[index == nb - 1 ? 0 : index + 1];
    vec3 line_dir = {{ep->x - sp->x, ep->y - sp->y, ep->z - sp->z}};
    normalize3(&line_dir);
    vec3 line_dir_p = cross3(&line_dir, &plane_norm);
    normalize3(&line_dir_p);
    orc_point far;
    far.x = p_proj.x + lambda * line_dir_p.v[0];
    far.y = p_proj.y + lambda * line_dir_p.v[1];
    far.z = p_proj.z + lambda * line_dir_p.v[2];
    far.w = 1.0f;
    int count = 0;
    for (int j = 0; j < nb; ++j) {
      const orc_point* a = &border[j];
      const orc_point* b = &border[j == nb - 1 ? 0 : j + 1];
      if (orc_segs_intersect(a, b, &p_proj, &far)) ++count;
    }
    odd += count % 2;
  }
  return odd >= 5; /* count >= count_for_intersect.size() / 2 */
}

/* the claim loop of postProcessPlanes (:1530-1556) and the rebuild of source_cloud (:1560-1566) */
int orc_reabsorb(const orc_point* cloud, size_t n, const float* coeffs, const orc_point* border, const size_t* bo, int n_planes,
                 float t, unsigned seed, int32_t* absorbed, size_t cap, size_t* plane_offsets, int32_t* remaining_idx,
                 size_t* n_remaining) {
  unsigned char* claimed = (unsigned char*)calloc(n ? n : 1, 1);
  if (!claimed) return -1;
  size_t at = 0;
  int rc = 0;
  for (int j = 0; j < n_planes && rc == 0; ++j) {
    plane_offsets[j] = at;
    const int nb = (int)(bo[j + 1] - bo[j]);
    if (nb <= 0) { rc = -1; break; }
    for (size_t i = 0; i < n; ++i) {
      if (orc_point_in_poly(&cloud[i], coeffs + 4 * j, border + bo[j], nb, t, seed)) {
        if (at >= cap) { rc = -1; break; }
        absorbed[at++] = (int32_t)i;
        claimed[i] = 1;
      }
    }
  }
  plane_offsets[n_planes] = at;
  size_t r = 0;
  for (size_t i = 0; i < n; ++i)
    if (!claimed[i]) remaining_idx[r++] = (int32_t)i;
  *n_remaining = r;
  free(claimed);
  return rc;
}
