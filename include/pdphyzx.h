/*
 * pdphyzx.h -- host side of pdphyzx-b200: the pdphyzx C API surface, kept in C, in front of
 * a CUDA batch back end (include/px_batch.h).
 *
 * This is a from-scratch implementation of the API that ctotheameron/pdphyzx exposes
 * (registerPdPhyzx -> PdPhyzxAPI{world, body}; reference src/includes/px_api.h:9-14,
 * px_world_api.h:25-36, px_body_api.h:18-30) with the same type names, field names and
 * argument meaning, so example-style game code compiles against it unchanged:
 *
 *     #include "pd_api.h"
 *     #define PDPHYZX_UNIT PDPHYZX_UNIT_MM      // optional, default metres
 *     #define PDPHYZX_IMPLEMENTATION            // in exactly one translation unit
 *     #include "pdphyzx.h"
 *
 * What lives here (host, runs once per body or per API call): shape construction (convex
 * hull, normals, box), mass properties, the body store with its activeIndices iteration
 * order, the fixed-timestep clock, the unit-scaled force/impulse/gravity setters.  All of it
 * reproduces the reference arithmetic bit for bit (fast rsqrt based rcp/div/sqrt, LUT
 * sin/cos, evaluation order) because the values feed the device step verbatim.
 *
 * What does NOT live here: the world step.  px->world->step() keeps the reference's clock
 * semantics on the host (src/px_world.c:422-445) and hands K = 5 x fired-steps substeps to
 * the device through the batch layer.  There is no CPU solver in this header and no
 * fallback: without the CUDA library the step fails loudly.
 */
#ifndef PDPHYZX_B200_H
#define PDPHYZX_B200_H

#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#error "pdphyzx.h is a C header (the API uses `new` as a member name, as the reference does)"
#endif

/* ---------------------------------------------------------------- units (px_unit.h:18-27) */
#define PDPHYZX_UNIT_MM 1000.0f
#define PDPHYZX_UNIT_CM 100.0f
#define PDPHYZX_UNIT_M 1.0f
#ifndef PDPHYZX_UNIT
#define PDPHYZX_UNIT PDPHYZX_UNIT_M
#endif
#define PDPHYZX_UNIT_SQ (PDPHYZX_UNIT * PDPHYZX_UNIT)

/* ------------------------------------------------------------- constants (px_math.h:22-34)
 * PX_2_PI / PX_H_PI / PX_EPSILON are deliberately unparenthesised products, as in the
 * reference: `a * PX_EPSILON * PX_EPSILON` style expressions must associate identically. */
#define PX_ONE_THIRD 0.333333333f
#define PX_EPSILON 0.0001f * PDPHYZX_UNIT
#define PX_PI 3.141592741f
#define PX_2_PI PX_PI * 2.0f
#define PX_H_PI PX_PI * 0.5f
#define PX_SIN_TABLE_SIZE 256
#define PX_SIN_TABLE_MASK (PX_SIN_TABLE_SIZE - 1)
#define PX_2PI_OVER_TABLE_SIZE (PX_2_PI / PX_SIN_TABLE_SIZE)

#define PX_VEC2_ARRAY_MAX_SIZE 32
#define PX_MAX_POLY_VERTEX_COUNT PX_VEC2_ARRAY_MAX_SIZE
#define PX_INVALID_WORLD_INDEX 255
/* The reference ships 64 (px_body_array.h:11) while its README promises 256-body worlds;
 * the index type is uint8_t with 255 as the invalid sentinel, so 255 is the real ceiling. */
#define PX_MAX_BODIES 255
#define PX_MAX_CONTACTS 2

/* ------------------------------------------------------------------ fast math (px_math.h) */

/* Quake-style inverse square root, one Newton step (px_math.h:50-64). */
static inline float pxFastRsqrt(float x) {
  union {
    float f;
    uint32_t u;
  } v = {.f = x};
  v.u = 0x5f3759dfu - (v.u >> 1);
  float y = v.f;
  return y * (1.5f - (x * 0.5f * y * y));
}
static inline float pxFastRcp(float x) {
  float r = pxFastRsqrt(x);
  return r * r;
}
static inline float pxFastSafeRcp(float x) { return x ? pxFastRcp(x) : 0; }
static inline float pxFastDiv(float num, float den) { return num * pxFastRcp(den); }
static inline float pxFastSqrt(float x) { return x * pxFastRsqrt(x); }
static inline float pxClamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline float pxMin(float a, float b) { return a < b ? a : b; }

/* 256-entry sine table, filled by registerPdPhyzx from the host libm (px_math.h:206-218).
 * One table per process: it is the table the device receives. */
extern float pxSinTable[PX_SIN_TABLE_SIZE];
void pxInitSinTable(void);

/* LUT sine with linear interpolation; the table index goes through the approximate divide
 * (px_math.h:226-241), so pxFastSin(1) = 0.83972 and not sinf(1). */
static inline float pxFastSin(float radians) {
  if ((radians = fmodf(radians, PX_2_PI)) < 0) {
    radians += PX_2_PI;
  }
  float pos = pxFastDiv(radians, PX_2PI_OVER_TABLE_SIZE);
  int i1 = (int)pos & PX_SIN_TABLE_MASK;
  int i2 = (i1 + 1) & PX_SIN_TABLE_MASK;
  float frac = pos - (int)pos;
  return pxSinTable[i1] * (1.0f - frac) + pxSinTable[i2] * frac;
}
static inline float pxFastCos(float radians) { return pxFastSin(radians + PX_H_PI); }

/* ----------------------------------------------------------------------- vectors (px_vec2.h) */
typedef struct {
  union {
    float v[2];
    struct {
      float x, y;
    };
  };
} PxVec2;

static inline PxVec2 pxVec2(float x, float y) { return (PxVec2){.x = x, .y = y}; }
static inline PxVec2 pxVec2Neg(PxVec2 a) { return pxVec2(-a.x, -a.y); }
static inline PxVec2 pxVec2Add(PxVec2 a, PxVec2 b) { return pxVec2(a.x + b.x, a.y + b.y); }
static inline PxVec2 pxVec2Sub(PxVec2 a, PxVec2 b) { return pxVec2(a.x - b.x, a.y - b.y); }
static inline PxVec2 pxVec2Addf(PxVec2 a, float s) { return pxVec2(a.x + s, a.y + s); }
static inline PxVec2 pxVec2Subf(PxVec2 a, float s) { return pxVec2(a.x - s, a.y - s); }
static inline PxVec2 pxVec2Multf(PxVec2 a, float s) { return pxVec2(a.x * s, a.y * s); }
static inline PxVec2 pxVec2Divf(PxVec2 a, float s) { return pxVec2(pxFastDiv(a.x, s), pxFastDiv(a.y, s)); }
static inline float pxVec2Dot(PxVec2 a, PxVec2 b) { return a.x * b.x + a.y * b.y; }
static inline float pxVec2Cross(PxVec2 a, PxVec2 b) { return a.x * b.y - a.y * b.x; }
static inline float pxVec2LenSqr(PxVec2 a) { return a.x * a.x + a.y * a.y; }
static inline float pxVec2Len(PxVec2 a) { return pxFastSqrt(pxVec2LenSqr(a)); }
/* leaves vectors shorter than PX_EPSILON untouched (px_vec2.h:459-472) */
static inline PxVec2 pxVec2Normalized(PxVec2 a) {
  float len = pxVec2Len(a);
  if (len <= PX_EPSILON) {
    return a;
  }
  float inv = pxFastRcp(len);
  return pxVec2(a.x * inv, a.y * inv);
}
static inline PxVec2 *pxVec2Normalize(PxVec2 *a) {
  *a = pxVec2Normalized(*a);
  return a;
}

typedef struct {
  PxVec2 items[2];
  uint8_t length;
} PxVec2Array2;

typedef struct {
  PxVec2 items[PX_VEC2_ARRAY_MAX_SIZE];
  uint8_t length;
} PxVec2Array;

/* -------------------------------------------------------------------- matrices (px_mat2.h) */
typedef struct {
  union {
    struct {
      float m00, m01;
      float m10, m11;
    };
    float m[2][2];
    float v[4];
  };
} PxMat2;

static inline PxMat2 pxMat2Identity(void) { return (PxMat2){.m00 = 1, .m01 = 0, .m10 = 0, .m11 = 1}; }
static inline PxMat2 pxMat2Orientation(float radians) {
  float c = pxFastCos(radians);
  float s = pxFastSin(radians);
  return (PxMat2){.m00 = c, .m01 = -s, .m10 = s, .m11 = c};
}
static inline PxMat2 pxMat2Transpose(PxMat2 a) {
  return (PxMat2){.m00 = a.m00, .m01 = a.m10, .m10 = a.m01, .m11 = a.m11};
}
static inline PxVec2 pxMat2MultVec2(PxMat2 a, PxVec2 v) {
  return pxVec2(a.m00 * v.x + a.m01 * v.y, a.m10 * v.x + a.m11 * v.y);
}

/* ------------------------------------------------------------------ colliders (px_collider.h) */
typedef enum { PX_CIRCLE, PX_POLYGON } PxColliderType;

typedef struct PxCircleData {
  float radius;
} PxCircleData;

typedef struct PxPolygonData {
  PxVec2Array vertices; /* hull, counter-clockwise, recentred on the centroid for dynamic bodies */
  PxVec2Array normals;  /* outward unit face normals, normals[i] belongs to edge i -> i+1 */
  float maxRadius;      /* farthest input vertex from the ORIGINAL origin (pre-recentring) */
} PxPolygonData;

typedef union {
  PxCircleData circle;
  PxPolygonData polygon;
} PxShapeData;

typedef struct {
  float mass, iMass;
  float momentOfInertia, iMomentOfInertia;
} PxMassData;

typedef struct {
  PxVec2 min;
  PxVec2 max;
} PxAABB;

typedef struct PxCollider PxCollider;
typedef PxMassData (*PxColliderComputeMassImpl)(PxCollider *collider, float density);
typedef void (*PxColliderUpdateAABB)(PxCollider *collider, PxAABB *outAABB, PxVec2 position);
typedef void (*PxColliderDestroyImpl)(PxCollider *collider);

struct PxCollider {
  PxShapeData shape;
  PxColliderType type;
  PxColliderComputeMassImpl computeMass;
  PxColliderUpdateAABB updateAABB;
  PxColliderDestroyImpl destroy;
};

PxCollider pxCircleColliderNew(float radius);
PxCollider pxBoxColliderNew(float width, float height);
PxCollider pxPolygonColliderNew(PxVec2 vertices[], uint8_t count);

/* ------------------------------------------------------------------------ bodies (px_body.h) */
typedef struct {
  float radius;
} PxCircle;

typedef struct __attribute__((aligned(4))) {
  PxVec2 vertices[PX_MAX_POLY_VERTEX_COUNT];
  size_t vertexCount;
} PxPolygon;

typedef struct {
  float width, height;
} PxBox;

/* Constructor argument: fill exactly one member; precedence box > circle > polygon
 * (px_body.c:16-25). */
typedef struct __attribute__((aligned(4))) {
  PxCircle circle;
  PxPolygon polygon;
  PxBox box;
} PxShape;

typedef enum {
  PX_BODY_STATIC = 0,
  PX_BODY_DYNAMIC = 1,
  PX_BODY_KINEMATIC = 2,
  PX_BODY_FLAG_VALID = (1 << 2),
  PX_BODY_FLAG_SLEEPING = (1 << 3),
  PX_BODY_TYPE_MASK = 0x03
} PxBodyFlags;

/* Public, caller-mutable (examples/draw/circle.c:43-45 pokes the material fields); field
 * order follows px_body.h:82-120. */
typedef struct __attribute__((aligned(4))) {
  uint8_t flags;
  uint8_t worldIndex;
  float sleepTime;
  float mass, iMass;
  PxVec2 position;
  PxVec2 velocity;
  float angularVelocity;
  float orientationAngle;
  PxVec2 force;
  float torque;
  float momentOfInertia, iMomentOfInertia;
  float restitution, staticFriction, dynamicFriction, density;
  PxMat2 orientation;
  PxAABB aabb;
  PxCollider collider;
} PxBody;

static inline bool pxBodyHasFlag(PxBody *b, PxBodyFlags f) { return (b->flags & f) != 0; }
static inline void pxBodySetFlag(PxBody *b, PxBodyFlags f, bool on) {
  b->flags = on ? (b->flags | f) : (b->flags & ~f);
}
static inline PxBodyFlags pxBodyGetType(PxBody *b) { return (PxBodyFlags)(b->flags & PX_BODY_TYPE_MASK); }
static inline bool pxBodyIsSleeping(PxBody *b) { return pxBodyHasFlag(b, PX_BODY_FLAG_SLEEPING); }
static inline bool pxBodyIsValid(PxBody *b) { return b != NULL && pxBodyHasFlag(b, PX_BODY_FLAG_VALID); }
static inline bool pxBodyIsStatic(PxBody *b) { return pxBodyGetType(b) == PX_BODY_STATIC; }
static inline bool pxBodyIsDynamic(PxBody *b) { return pxBodyGetType(b) == PX_BODY_DYNAMIC; }

PxBody pxBodyNew(PxShape shape, PxBodyFlags type, float density, PxVec2 position);
void pxBodySetPosition(PxBody *body, float x, float y);
void pxBodyMoveBy(PxBody *body, PxVec2 distance);
void pxBodySetOrientation(PxBody *body, float radians);
void pxBodyRotate(PxBody *body, float radians);
void pxBodyApplyForce(PxBody *body, PxVec2 force, PxVec2 contact);
void pxBodyApplyImpulse(PxBody *body, PxVec2 impulse, PxVec2 contact);
void pxBodyClearForces(PxBody *body);
void pxBodyWakeUp(PxBody *body);

/* -------------------------------------------------------------- body store (px_body_array.h) */
typedef struct __attribute__((aligned(4))) {
  PxBody items[PX_MAX_BODIES];
  uint8_t length;     /* active bodies */
  uint8_t freedCount; /* reusable slots on the freed stack */
  uint8_t activeIndices[PX_MAX_BODIES]; /* iteration order; the dynamic array keeps the
                                           x-sorted order the step leaves behind */
  uint8_t freedIndices[PX_MAX_BODIES];
} PxBodyArray;

#define pxBodyArrayEach(array, body)                                                              \
  for (uint8_t _i = 0; _i < (array)->length; _i++)                                                \
    for (PxBody *body = &(array)->items[(array)->activeIndices[_i]]; body; body = NULL)

PxBody *pxBodyArrayAdd(PxBodyArray *array, PxBody *body);
void pxBodyArrayRemove(PxBodyArray *array, uint8_t index);
void pxBodyArraySortByAxis(PxBodyArray *bodies);

/* ------------------------------------------------------------------------ clock (px_clock.h) */
typedef struct {
  uint32_t lastUpdateMs;
  float accumulator;
  float fixedTimestep;
  uint8_t targetFps;
  uint8_t maxStepsPerFrame;
  bool isFirstUpdate;
} PxClock;

/* ------------------------------------------------------------------------ world (px_world.h) */
#define PX_SUBSTEPS_PER_STEP 5 /* px_world.c:427 */

/* Contact statistics of the last device substep.  The reference keeps its manifold pool in
 * the world (px_world.h:34) and callers only ever read `contacts.length` (drawDebug,
 * px_world.c:495-496); the manifolds themselves live and die on the device here. */
typedef struct {
  uint16_t length;   /* manifolds kept by the last substep */
  uint16_t sleeping; /* dynamic bodies asleep after the last substep */
  uint32_t overflow; /* PX_BATCH_OVERFLOW_* bits; 0 = fine */
} PxContactStats;

typedef struct PxWorld {
  PxBodyArray staticBodies;
  PxBodyArray dynamicBodies;
  PxContactStats contacts;

  uint8_t iterations;
  PxVec2 gravity;

  float linearSleepThresholdSq;
  float angularSleepThreshold;
  float timeToSleep;

  PxClock clock;

  void *batch; /* lazily created 1-world PxBatch behind px->world->step; NULL until stepped */
} PxWorld;

PxWorld *pxWorldNew(uint8_t targetFps);
void pxWorldFree(PxWorld *world);
void pxWorldSetGravity(PxWorld *world, float x, float y);
void pxWorldSetIterations(PxWorld *world, uint8_t iterations);
PxBody *pxWorldNewStaticBody(PxWorld *world, PxShape shape, PxVec2 position);
PxBody *pxWorldNewDynamicBody(PxWorld *world, PxShape shape, float density, PxVec2 position);
void pxWorldFreeBody(PxWorld *world, PxBody *body);
bool pxWorldStep(PxWorld *world);
void pxWorldDrawDebug(PxWorld *world);
/* substep dt the step uses: pxFastDiv(fixedTimestep, 5) (px_world.c:432) */
float pxWorldSubstepDt(const PxWorld *world);

/* -------------------------------------------------------------------------------- API tables */
typedef struct {
  PxWorld *(*new)(uint8_t targetFps);
  void (*free)(PxWorld *world);
  void (*setIterations)(PxWorld *world, uint8_t iterations);
  void (*setGravity)(PxWorld *world, float x, float y);
  PxBody *(*newStaticBody)(PxWorld *world, PxShape shape, PxVec2 position);
  PxBody *(*newDynamicBody)(PxWorld *world, PxShape shape, float density, PxVec2 position);
  void (*freeBody)(PxWorld *world, PxBody *body);
  bool (*step)(PxWorld *world);
  void (*drawDebug)(PxWorld *world);
} PxWorldAPI;

typedef struct {
  bool (*isValid)(PxBody *body);
  void (*setPosition)(PxBody *body, float x, float y);
  void (*setOrientation)(PxBody *body, float radians);
  void (*moveBy)(PxBody *body, PxVec2 distance);
  void (*rotate)(PxBody *body, float radians);
  void (*applyForce)(PxBody *body, float forceX, float forceY, float contactX, float contactY);
  void (*applyImpulse)(PxBody *body, float impulseX, float impulseY, float contactX, float contactY);
} PxBodyAPI;

typedef struct {
  const PxBodyAPI *body;
  const PxWorldAPI *world;
} PdPhyzxAPI;

#ifdef PD_API_HOST_STUB_H
#define PDPHYZX_HAVE_PD_API 1
#elif defined(TARGET_EXTENSION) || defined(PLAYDATEAPI_H) || defined(pdext_api_h)
#define PDPHYZX_HAVE_PD_API 1
#endif
#ifndef PDPHYZX_HAVE_PD_API
typedef struct PlaydateAPI PlaydateAPI; /* include "pd_api.h" first for the full type */
#endif

PdPhyzxAPI *registerPdPhyzx(PlaydateAPI *playdate);

/* ==================================================================== implementation ==== */
#ifdef PDPHYZX_IMPLEMENTATION

#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "px_batch.h"

#ifndef PDPHYZX_HAVE_PD_API
#error "define PDPHYZX_IMPLEMENTATION only after including pd_api.h (the SDK's or oracle/stub/pd_api.h)"
#endif

PlaydateAPI *pd = NULL;
float pxSinTable[PX_SIN_TABLE_SIZE];

#ifndef pxalloc
#define pxalloc(n) pd->system->realloc(NULL, (n))
#endif
#ifndef pxfree
#define pxfree(p) pd->system->realloc((p), 0)
#endif

void pxInitSinTable(void) {
  for (int i = 0; i < PX_SIN_TABLE_SIZE; i++) {
    pxSinTable[i] = sinf((i * PX_2_PI) / PX_SIN_TABLE_SIZE);
  }
}

/* ---- circle ------------------------------------------------------------------------------ */

/* mass = pi r^2 rho, I = mass r^2 (px_circle.c:13-34) */
static PxMassData pxCircleComputeMass(PxCollider *c, float density) {
  if (!c) {
    return (PxMassData){0};
  }
  float r = c->shape.circle.radius;
  float rr = r * r;
  float mass = PX_PI * rr * density;
  float inertia = mass * rr;
  return (PxMassData){mass, pxFastSafeRcp(mass), inertia, pxFastSafeRcp(inertia)};
}

/* AABB = centre -/+ radius (px_circle.c:43-48) */
static void pxCircleUpdateAABB(PxCollider *c, PxAABB *out, PxVec2 position) {
  float r = c->shape.circle.radius;
  out->min = pxVec2Subf(position, r);
  out->max = pxVec2Addf(position, r);
}

PxCollider pxCircleColliderNew(float radius) {
  PxCollider c;
  memset(&c, 0, sizeof(c));
  c.type = PX_CIRCLE;
  c.shape.circle.radius = radius;
  c.computeMass = &pxCircleComputeMass;
  c.updateAABB = &pxCircleUpdateAABB;
  return c;
}

/* ---- polygon ----------------------------------------------------------------------------- */

/*
 * Convex hull by gift wrapping from the right-most (then lowest-y) input point, keeping the
 * farther of collinear candidates; then maxRadius over the hull points as given (before any
 * recentring) and the outward edge normals (px_polygon.c:15-126).
 */
static PxPolygonData pxPolygonData(uint8_t count, const PxVec2 *pts) {
  PxPolygonData poly;
  memset(&poly, 0, sizeof(poly));
  if (count < 3 || count > PX_MAX_POLY_VERTEX_COUNT) {
    return poly; /* the reference asserts; a zero-vertex polygon makes an invalid body here */
  }

  uint8_t start = 0;
  for (uint8_t i = 1; i < count; i++) {
    if (pts[i].x > pts[start].x || (pts[i].x == pts[start].x && pts[i].y < pts[start].y)) {
      start = i;
    }
  }

  uint8_t hull[PX_MAX_POLY_VERTEX_COUNT];
  uint8_t n = 0;
  uint8_t cur = start;
  for (;;) {
    hull[n] = cur;
    uint8_t next = 0;
    for (uint8_t i = 1; i < count; i++) {
      if (next == cur) {
        next = i;
        continue;
      }
      PxVec2 e1 = pxVec2Sub(pts[next], pts[hull[n]]);
      PxVec2 e2 = pxVec2Sub(pts[i], pts[hull[n]]);
      float c = pxVec2Cross(e1, e2);
      if (c < 0 || (c == 0 && pxVec2LenSqr(e2) > pxVec2LenSqr(e1))) {
        next = i;
      }
    }
    n++;
    cur = next;
    if (next == start || n >= PX_MAX_POLY_VERTEX_COUNT) {
      break;
    }
  }
  poly.vertices.length = n;
  poly.normals.length = n;

  poly.maxRadius = 0;
  for (uint8_t i = 0; i < n; i++) {
    PxVec2 v = pts[hull[i]];
    poly.vertices.items[i] = v;
    float d2 = pxVec2LenSqr(v);
    if (d2 > poly.maxRadius * poly.maxRadius) {
      poly.maxRadius = pxFastSqrt(d2);
    }
  }
  for (uint8_t i = 0; i < n; i++) {
    uint8_t j = (uint8_t)(i + 1 < n ? i + 1 : 0);
    PxVec2 edge = pxVec2Sub(poly.vertices.items[j], poly.vertices.items[i]);
    poly.normals.items[i] = pxVec2Normalized(pxVec2(edge.y, -edge.x));
  }
  return poly;
}

/*
 * Area, centroid (then the hull is shifted so the centroid is the origin), polygon inertia
 * about it, mass = rho * area (px_polygon.c:145-252).
 */
static PxMassData pxPolygonComputeMass(PxCollider *c, float density) {
  if (!c) {
    return (PxMassData){0};
  }
  PxVec2Array *vs = &c->shape.polygon.vertices;
  uint8_t n = vs->length;

  float area = 0;
  PxVec2 centroid = pxVec2(0, 0);
  for (uint8_t i = 0; i < n; i++) {
    uint8_t j = (uint8_t)(i + 1 < n ? i + 1 : 0);
    PxVec2 a = vs->items[i], b = vs->items[j];
    float tri = 0.5f * pxVec2Cross(a, b);
    area += tri;
    PxVec2 w = pxVec2Multf(pxVec2Add(a, b), tri * PX_ONE_THIRD);
    centroid.x += w.x;
    centroid.y += w.y;
  }
  float invArea = pxFastRcp(area);
  centroid.x *= invArea;
  centroid.y *= invArea;
  for (uint8_t i = 0; i < n; i++) {
    vs->items[i].x -= centroid.x;
    vs->items[i].y -= centroid.y;
  }

  float inertia = 0;
  for (uint8_t i = 0; i < n; i++) {
    uint8_t j = (uint8_t)(i + 1 < n ? i + 1 : 0);
    PxVec2 a = vs->items[i], b = vs->items[j];
    float cross = pxVec2Cross(a, b);
    float sq = a.x * a.x + a.x * b.x + b.x * b.x + a.y * a.y + a.y * b.y + b.y * b.y;
    inertia += cross * sq;
  }
  inertia *= pxFastDiv(density, 12.0f);
  float mass = density * area;
  return (PxMassData){mass, pxFastSafeRcp(mass), inertia, pxFastSafeRcp(inertia)};
}

/* AABB = centre -/+ maxRadius, orientation-independent (px_polygon.c:261-266) */
static void pxPolygonUpdateAABB(PxCollider *c, PxAABB *out, PxVec2 position) {
  float r = c->shape.polygon.maxRadius;
  out->min = pxVec2Subf(position, r);
  out->max = pxVec2Addf(position, r);
}

PxCollider pxPolygonColliderNew(PxVec2 vertices[], uint8_t count) {
  PxCollider c;
  memset(&c, 0, sizeof(c));
  c.type = PX_POLYGON;
  c.shape.polygon = pxPolygonData(count, vertices);
  c.computeMass = &pxPolygonComputeMass;
  c.updateAABB = &pxPolygonUpdateAABB;
  return c;
}

/* axis-aligned box, CCW from (-w/2,-h/2), exact unit normals (px_polygon.c:278-324) */
PxCollider pxBoxColliderNew(float width, float height) {
  PxCollider c;
  memset(&c, 0, sizeof(c));
  if (width <= 0.0f || height <= 0.0f) {
    return c;
  }
  float hw = width * 0.5f, hh = height * 0.5f;
  PxPolygonData *p = &c.shape.polygon;
  p->vertices.length = p->normals.length = 4;
  p->vertices.items[0] = pxVec2(-hw, -hh);
  p->vertices.items[1] = pxVec2(hw, -hh);
  p->vertices.items[2] = pxVec2(hw, hh);
  p->vertices.items[3] = pxVec2(-hw, hh);
  p->normals.items[0] = pxVec2(0, -1);
  p->normals.items[1] = pxVec2(1, 0);
  p->normals.items[2] = pxVec2(0, 1);
  p->normals.items[3] = pxVec2(-1, 0);
  p->maxRadius = pxFastSqrt(hw * hw + hh * hh);
  c.type = PX_POLYGON;
  c.computeMass = &pxPolygonComputeMass;
  c.updateAABB = &pxPolygonUpdateAABB;
  return c;
}

/* ---- body -------------------------------------------------------------------------------- */

/* px_body.c:12-83: collider by precedence box > circle > polygon; AABB before the mass pass
 * (which recentres polygon vertices); density is per unit area in metres, hence / UNIT^2. */
PxBody pxBodyNew(PxShape shape, PxBodyFlags type, float density, PxVec2 position) {
  PxBody body;
  memset(&body, 0, sizeof(body));

  PxCollider collider;
  if (shape.box.width != 0 && shape.box.height != 0) {
    collider = pxBoxColliderNew(shape.box.width, shape.box.height);
  } else if (shape.circle.radius != 0) {
    collider = pxCircleColliderNew(shape.circle.radius);
  } else if (shape.polygon.vertexCount != 0) {
    collider = pxPolygonColliderNew(shape.polygon.vertices, (uint8_t)shape.polygon.vertexCount);
  } else {
    return body;
  }
  if (!collider.updateAABB || (collider.type == PX_POLYGON && collider.shape.polygon.vertices.length == 0)) {
    return body; /* rejected box / polygon: invalid body (the reference would crash or assert) */
  }

  PxAABB aabb = {0};
  collider.updateAABB(&collider, &aabb, position);

  PxMassData md = {0};
  if (type == PX_BODY_DYNAMIC) {
    if (density <= 0) {
      return body;
    }
    md = collider.computeMass(&collider, pxFastDiv(density, PDPHYZX_UNIT_SQ));
  }

  body.flags = (uint8_t)(PX_BODY_FLAG_VALID | (type & PX_BODY_TYPE_MASK));
  body.worldIndex = PX_INVALID_WORLD_INDEX;
  body.aabb = aabb;
  body.collider = collider;
  body.density = density;
  body.position = position;
  body.dynamicFriction = 0.3f;
  body.staticFriction = 0.5f;
  body.restitution = 0.2f;
  body.orientation = pxMat2Identity();
  body.mass = md.mass;
  body.iMass = md.iMass;
  body.momentOfInertia = md.momentOfInertia;
  body.iMomentOfInertia = md.iMomentOfInertia;
  return body;
}

static void pxBodyRefreshAABB(PxBody *body) {
  body->collider.updateAABB(&body->collider, &body->aabb, body->position);
}

void pxBodySetPosition(PxBody *body, float x, float y) {
  if (!body) return;
  body->position = pxVec2(x, y);
  pxBodyRefreshAABB(body);
}

void pxBodyMoveBy(PxBody *body, PxVec2 d) {
  if (!body) return;
  body->position.x += d.x;
  body->position.y += d.y;
  pxBodyRefreshAABB(body);
}

/* angle wrapped into [0, 2pi) with fmodf, matrix from the LUT (px_body.c:104-116) */
void pxBodySetOrientation(PxBody *body, float radians) {
  if (!body) return;
  if ((radians = fmodf(radians, PX_2_PI)) < 0) {
    radians += PX_2_PI;
  }
  body->orientationAngle = radians;
  body->orientation = pxMat2Orientation(radians);
}

void pxBodyRotate(PxBody *body, float radians) {
  if (!body) return;
  pxBodySetOrientation(body, body->orientationAngle + radians);
}

/* `contact` is used as the lever arm itself, not contact - position (px_body.c:147-165) */
void pxBodyApplyImpulse(PxBody *body, PxVec2 impulse, PxVec2 contact) {
  if (!body || !pxBodyIsDynamic(body)) return;
  body->velocity.x += impulse.x * body->iMass;
  body->velocity.y += impulse.y * body->iMass;
  body->angularVelocity += pxVec2Cross(contact, impulse) * body->iMomentOfInertia;
}

void pxBodyApplyForce(PxBody *body, PxVec2 force, PxVec2 contact) {
  if (!body || !pxBodyIsDynamic(body)) return;
  body->force.x += force.x;
  body->force.y += force.y;
  body->torque += pxVec2Cross(contact, force);
}

void pxBodyClearForces(PxBody *body) {
  body->force = pxVec2(0, 0);
  body->torque = 0;
}

void pxBodyWakeUp(PxBody *body) {
  if (!body) return;
  body->sleepTime = 0.0f;
  pxBodySetFlag(body, PX_BODY_FLAG_SLEEPING, false);
}

/* ---- body store -------------------------------------------------------------------------- */

/* freed slots are reused first (LIFO), otherwise slot = length; appended to the iteration
 * order (px_body_array.c:7-59).  Note slot = length is only collision-free while nothing has
 * been removed without reuse -- same as the reference. */
PxBody *pxBodyArrayAdd(PxBodyArray *array, PxBody *body) {
  if (!pxBodyIsValid(body)) {
    return NULL;
  }
  uint8_t slot;
  if (array->freedCount > 0) {
    slot = array->freedIndices[--array->freedCount];
  } else if (array->length < PX_MAX_BODIES) {
    slot = array->length;
  } else {
    return NULL;
  }
  array->activeIndices[array->length++] = slot;
  body->worldIndex = slot;
  array->items[slot] = *body;
  return &array->items[slot];
}

/* swap-remove from the iteration order; `index >= length` is a silent no-op even for a live
 * slot (px_body_array.c:61-79, quirk kept) */
void pxBodyArrayRemove(PxBodyArray *array, uint8_t index) {
  if (index >= array->length || array->freedCount >= PX_MAX_BODIES) {
    return;
  }
  array->freedIndices[array->freedCount++] = index;
  for (uint8_t i = 0; i < array->length; i++) {
    if (array->activeIndices[i] == index) {
      array->activeIndices[i] = array->activeIndices[--array->length];
      break;
    }
  }
}

/* stable insertion sort of the iteration order by aabb.min.x (px_body_array.c:81-102) */
void pxBodyArraySortByAxis(PxBodyArray *bodies) {
  for (int i = 1; i < bodies->length; i++) {
    uint8_t slot = bodies->activeIndices[i];
    float key = bodies->items[slot].aabb.min.x;
    int j = i - 1;
    while (j >= 0 && bodies->items[bodies->activeIndices[j]].aabb.min.x > key) {
      bodies->activeIndices[j + 1] = bodies->activeIndices[j];
      j--;
    }
    bodies->activeIndices[j + 1] = slot;
  }
}

/* ---- clock (px_clock.c:10-52) ------------------------------------------------------------ */

static void pxClockInit(PxClock *clock, uint8_t targetFps) {
  clock->maxStepsPerFrame = 3;
  clock->targetFps = (uint8_t)pxClamp(targetFps <= 0 ? 50 : targetFps, 1, 50);
  clock->fixedTimestep = pxFastRcp(clock->targetFps);
  clock->lastUpdateMs = 0;
  clock->accumulator = 0;
  clock->isFirstUpdate = true;
}

static void pxClockBeginFrame(PxClock *clock) {
  uint32_t now = pd->system->getCurrentTimeMilliseconds();
  if (clock->isFirstUpdate) {
    clock->lastUpdateMs = now;
    clock->isFirstUpdate = false;
    return;
  }
  uint32_t elapsed = now - clock->lastUpdateMs;
  clock->accumulator += pxClamp(pxFastDiv(elapsed, 1000), 0, 0.1f);
  clock->lastUpdateMs = now;
}

/* ---- world ------------------------------------------------------------------------------- */

PxWorld *pxWorldNew(uint8_t targetFps) {
  PxWorld *world = (PxWorld *)pxalloc(sizeof(PxWorld));
  if (!world) {
    fprintf(stderr, "pdphyzx: failed to allocate a PxWorld\n");
    abort();
  }
  memset(world, 0, sizeof(*world));
  world->iterations = 8;                                  /* px_world.c:14 */
  world->linearSleepThresholdSq = 0.001f * PDPHYZX_UNIT_SQ; /* px_world.c:15 */
  world->angularSleepThreshold = 0.01f;                   /* px_world.c:16 */
  world->timeToSleep = 0.5f;                              /* px_world.c:17 */
  pxClockInit(&world->clock, targetFps);
  return world;
}

void pxWorldFree(PxWorld *world) {
  if (!world) return;
  if (world->batch) {
    pxBatchFree((PxBatch *)world->batch);
  }
  pxfree(world);
}

void pxWorldSetGravity(PxWorld *world, float x, float y) {
  if (!world) return;
  world->gravity = pxVec2Multf(pxVec2(x, y), PDPHYZX_UNIT);
}

void pxWorldSetIterations(PxWorld *world, uint8_t iterations) {
  if (!world) return;
  world->iterations = iterations;
}

PxBody *pxWorldNewStaticBody(PxWorld *world, PxShape shape, PxVec2 position) {
  PxBody body = pxBodyNew(shape, PX_BODY_STATIC, 0, position);
  PxBody *stored = pxBodyArrayAdd(&world->staticBodies, &body);
  pxBodyArraySortByAxis(&world->staticBodies); /* px_world.c:98 */
  return stored; /* the device copy follows at the next step: the facade's batch has headroom */
}

PxBody *pxWorldNewDynamicBody(PxWorld *world, PxShape shape, float density, PxVec2 position) {
  PxBody body = pxBodyNew(shape, PX_BODY_DYNAMIC, density, position);
  PxBody *stored = pxBodyArrayAdd(&world->dynamicBodies, &body);
  return stored;
}

void pxWorldFreeBody(PxWorld *world, PxBody *body) {
  if (!world || !body) return;
  PxBodyArray *arr = body->density > 0 ? &world->dynamicBodies : &world->staticBodies;
  pxBodyArrayRemove(arr, body->worldIndex);
}

float pxWorldSubstepDt(const PxWorld *world) {
  return pxFastDiv(world->clock.fixedTimestep, PX_SUBSTEPS_PER_STEP);
}

/*
 * px->world->step: the reference's accumulator loop (px_world.c:422-445) decides how many
 * fixed steps fire (0..3); the substeps themselves run on the device.  Host-side edits made
 * between calls (positions, velocities, forces, materials, gravity ...) are carried over by
 * re-uploading the world before the substeps and reading it back after them.
 */
bool pxWorldStep(PxWorld *world) {
  PxClock *clock = &world->clock;
  pxClockBeginFrame(clock);
  uint8_t fired = 0;
  while (clock->accumulator >= clock->fixedTimestep && fired < clock->maxStepsPerFrame) {
    clock->accumulator -= clock->fixedTimestep;
    fired++;
  }
  if (!fired) {
    return false;
  }
  PxWorld *const one[1] = {world};
  if (world->batch && pxBatchUpload((PxBatch *)world->batch, one, 0, 1) != 0) {
    /* the world outgrew the batch's capacities (it was made with generous headroom): make a bigger one */
    pxBatchFree((PxBatch *)world->batch);
    world->batch = NULL;
  }
  if (!world->batch) {
    /* One batch for the life of the world: bodies come and go (examples/draw spawns and frees
     * circles while running), so the capacities leave room -- every dynamic slot, twice the
     * statics and vertices present now -- and add/free never re-creates device buffers. */
    PxBatchDesc desc = pxBatchDescDefault(PDPHYZX_UNIT);
    if (pxBatchFitDesc(one, 1, &desc) != 0) {
      fprintf(stderr, "pdphyzx: %s\n", pxBatchLastError());
      abort();
    }
    desc.maxDynamic = PX_MAX_BODIES;
    desc.maxStatic = desc.maxStatic * 2 > 16 ? desc.maxStatic * 2 : 16;
    if (desc.maxStatic > 64) desc.maxStatic = 64 > world->staticBodies.length ? 64 : world->staticBodies.length;
    desc.maxVertices = desc.maxVertices * 2 > 2048 ? desc.maxVertices * 2 : 2048;
    world->batch = pxBatchNew(one, 1, &desc);
    if (!world->batch) {
      fprintf(stderr, "pdphyzx: px->world->step needs the CUDA batch back end: %s\n", pxBatchLastError());
      abort();
    }
  }
  int rc = pxBatchStep((PxBatch *)world->batch, (uint32_t)fired * PX_SUBSTEPS_PER_STEP);
  if (rc == 0) {
    rc = pxBatchReadback((PxBatch *)world->batch, one, 0, 1, 0);
  }
  if (rc < 0) {
    fprintf(stderr, "pdphyzx: device step failed: %s\n", pxBatchLastError());
    abort();
  }
  return true;
}

/* AABB rectangles, a square on sleepers, velocity whiskers, "Contacts/Sleeping" text
 * (px_world.c:451-499) */
void pxWorldDrawDebug(PxWorld *world) {
  uint8_t sleeping = 0;
  pxBodyArrayEach(&world->staticBodies, body) {
    PxAABB a = body->aabb;
    pd->graphics->drawRect(a.min.x, a.min.y, a.max.x - a.min.x, a.max.y - a.min.y, kColorXOR);
  }
  pxBodyArrayEach(&world->dynamicBodies, body) {
    PxAABB a = body->aabb;
    pd->graphics->drawRect(a.min.x, a.min.y, a.max.x - a.min.x, a.max.y - a.min.y, kColorXOR);
    float cx = (a.min.x + a.max.x) / 2.0f, cy = (a.min.y + a.max.y) / 2.0f;
    if (pxBodyIsSleeping(body)) {
      sleeping++;
      pd->graphics->drawRect(cx - 4, cy - 4, 8, 8, kColorXOR);
    }
    if (pxVec2Len(body->velocity) > 0.01f) {
      pd->graphics->drawLine(cx, cy, cx + body->velocity.x * 10, cy + body->velocity.y * 10, 1, kColorBlack);
    }
  }
  char *text = NULL;
  int len = pd->system->formatString(&text, "Contacts: %d  Sleeping: %d ", world->contacts.length, sleeping);
  pd->graphics->drawText(text, len, kASCIIEncoding, 5, 20);
}

/* ---- API tables (px_api.c, px_world_api.c, px_body_api.c) --------------------------------- */

static void pxBodyApiApplyForce(PxBody *body, float fx, float fy, float cx, float cy) {
  pxBodyApplyForce(body, pxVec2Multf(pxVec2(fx, fy), PDPHYZX_UNIT), pxVec2(cx, cy));
  pxBodyWakeUp(body);
}

static void pxBodyApiApplyImpulse(PxBody *body, float ix, float iy, float cx, float cy) {
  pxBodyApplyImpulse(body, pxVec2Multf(pxVec2(ix, iy), PDPHYZX_UNIT), pxVec2(cx, cy));
  pxBodyWakeUp(body);
}

static bool pxBodyApiIsValid(PxBody *body) { return pxBodyIsValid(body); }

const PxWorldAPI *px_world = NULL;
const PxBodyAPI *px_body = NULL;

PdPhyzxAPI *registerPdPhyzx(PlaydateAPI *playdate) {
  pd = playdate;
  pxInitSinTable();

  static PxWorldAPI worldApi;
  static PxBodyAPI bodyApi;
  static PdPhyzxAPI api;

  worldApi.new = &pxWorldNew;
  worldApi.free = &pxWorldFree;
  worldApi.setIterations = &pxWorldSetIterations;
  worldApi.setGravity = &pxWorldSetGravity;
  worldApi.newStaticBody = &pxWorldNewStaticBody;
  worldApi.newDynamicBody = &pxWorldNewDynamicBody;
  worldApi.freeBody = &pxWorldFreeBody;
  worldApi.step = &pxWorldStep;
  worldApi.drawDebug = &pxWorldDrawDebug;

  bodyApi.isValid = &pxBodyApiIsValid;
  bodyApi.setPosition = &pxBodySetPosition;
  bodyApi.setOrientation = &pxBodySetOrientation;
  bodyApi.moveBy = &pxBodyMoveBy;
  bodyApi.rotate = &pxBodyRotate;
  bodyApi.applyForce = &pxBodyApiApplyForce;
  bodyApi.applyImpulse = &pxBodyApiApplyImpulse;

  px_world = &worldApi;
  px_body = &bodyApi;
  api.world = &worldApi;
  api.body = &bodyApi;
  return &api;
}

#endif /* PDPHYZX_IMPLEMENTATION */
#endif /* PDPHYZX_B200_H */
