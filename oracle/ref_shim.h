// ref_shim.h — the smallest stand-ins for the PCL / Eigen types that the reference's isPointInPoly family
// (Dialog/PlaneDetect.h) touches, so that THOSE FUNCTIONS' OWN SOURCE can be compiled here as a checker
// (oracle/build_ref.py -> oracle/_ref/libdialog_ref.so).  TEST INFRASTRUCTURE ONLY.
//
// Nothing in this file comes from the reference, PCL or Eigen; it restates only the arithmetic of the few
// Eigen members used (same CHOICES as pr_oracle.h: 3-coefficient reductions as e0 + (e1 + e2); normalize()
// divides by sqrt(squaredNorm) when it is > 0).  libm / CRT calls made by the reference are redirected:
// srand / rand -> the MSVC CRT generator, time(0) -> a settable value, pow(a, 0.5f) -> sqrtf(a).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace pcl {
struct PointXYZ { float x, y, z, w; };
struct ModelCoefficients { std::vector<float> values; };
template <class T> struct PointCloud {
  std::vector<T> points;
  typedef PointCloud* Ptr;
  size_t size() const { return points.size(); }
};
}  // namespace pcl

namespace Eigen {
template <int N> struct Vec {
  float c[N];
  struct Comma {
    Vec* v; int at;
    Comma& operator,(float s) { v->c[at++] = s; return *this; }
  };
  Comma operator<<(float s) { c[0] = s; return Comma{this, 1}; }
  float& operator[](int i) { return c[i]; }
  float operator[](int i) const { return c[i]; }
  float& operator()(int i) { return c[i]; }
  float operator()(int i) const { return c[i]; }
  float dot(const Vec& o) const {
    if (N == 3) return c[0] * o.c[0] + (c[1] * o.c[1] + c[2] * o.c[2]);
    return (c[0] * o.c[0] + c[1] * o.c[1]) + (c[2] * o.c[2] + c[3 % N] * o.c[3 % N]);
  }
  void normalize() {
    const float z = dot(*this);
    if (z > 0.0f) { const float n = std::sqrt(z); for (int i = 0; i < N; ++i) c[i] /= n; }
  }
  Vec cross(const Vec& o) const {
    Vec r;
    r.c[0] = c[1] * o.c[2] - c[2] * o.c[1];
    r.c[1] = c[2] * o.c[0] - c[0] * o.c[2];
    r.c[2] = c[0] * o.c[1] - c[1] * o.c[0];
    return r;
  }
};
typedef Vec<3> Vector3f;
typedef Vec<4> Vector4f;
}  // namespace Eigen

typedef pcl::PointXYZ PointT;
typedef pcl::PointCloud<PointT> PointCloudT;
using namespace std;

struct Plane {  // the members of Dialog/HeaderFile.h:81-88 that the extracted functions read
  PointCloudT::Ptr border;
  PointCloudT::Ptr points_set;
  pcl::ModelCoefficients coeff;
};

static float T_dist_point_plane = 0.1f;  // Dialog/PlaneDetect.h:88

// CRT redirections (the reference seeds with srand(time(0)) at every isPointInPoly call)
static unsigned ref_time_value = 0;
static unsigned long ref_hold = 1;
static inline unsigned ref_time() { return ref_time_value; }
static inline void ref_srand(unsigned s) { ref_hold = s; }
static inline int ref_rand() {
  ref_hold = (ref_hold * 214013ul + 2531011ul) & 0xFFFFFFFFul;
  return (int)((ref_hold >> 16) & 0x7fff);
}
static inline float ref_pow(float a, float) { return std::sqrt(a); }
#define srand ref_srand
#define rand ref_rand
#define time(x) ref_time()
#define pow ref_pow

bool isBothLineSegsIntersect(pcl::PointXYZ& pa, pcl::PointXYZ& pb, pcl::PointXYZ& pc, pcl::PointXYZ& pd, pcl::PointXYZ& p_inter);
void getInfoBetPointAndPlane(pcl::PointXYZ& p, Eigen::Vector4f& plane_param, float& dist, pcl::PointXYZ& p_proj);
bool isPointInPoly(pcl::PointXYZ& p, Plane& plane);
