/* pt_dev_scene.h -- the scene as the CUDA kernels read it.
 *
 * The reference's shader re-derives per-object constants for every ray: RotationMatrix() with six sin/cos per
 * rotated primitive (shader.comp:339,371,636), lens cap geometry (shader.comp:423-431), bounding-sphere radii
 * (shader.comp:887,899-905), the camera basis and the camera lens (shader.comp:1456,1409-1418) -- all functions
 * of the uniform block and the push constants only.  libpt_cuda evaluates them ONCE on the host
 * (pt_prepare.cpp), with the same pt_math.h functions and the same fp32 operation order the shader uses, so the
 * hoisted values are bit-identical to what a per-ray evaluation would produce, and passes the result to the
 * kernel as a __grid_constant__ parameter: it lives in the constant bank, where warp-uniform reads are free
 * operands of the FP32 pipe.
 *
 * Plain C structs of floats/ints; shared by g++ (host), nvcc and NVRTC.
 */
#ifndef PT_DEV_SCENE_H
#define PT_DEV_SCENE_H

/* Capacity = whatever fits the reference's uniform block (1024 object floats, 768 SDF floats: host:39-40), i.e. up to
 * 170 spheres / 204 planes / 93 boxes / 85 lenses / 64 cyclides, and 128 SDFs (sdfs[768]; four 32-bit masks).
 * The records of all types sit back to back in one pool, in the shader's type order; the worst case for the pool is
 * a scene made of boxes only (20 pool floats per 11 uniform-block floats). */
#define PT_DEV_MAX_SDFS 128
#define PT_DEV_POOL_FLOATS (((1024 / 11) * 20) + 24 + PT_DEV_MAX_SDFS * 8)
#define PT_DEV_MAX_LIGHT_SLOTS 65 /* lightIDs[0..numLights] inclusive: r == 1.0 reads one past */

typedef struct PtDevSphere { /* shader.comp:59-64, 289-317 */
    float px, py, pz, radius;
    float r2;            /* radius * radius */
    float materialID;    /* float(int(raw) - 1) */
    float lightID;
    float pad;
} PtDevSphere;

typedef struct PtDevPlane { /* shader.comp:66-70, 319-335: only pos.y matters */
    float py, materialID, lightID, pad;
} PtDevPlane;

typedef struct PtDevBox { /* shader.comp:72-78, 337-364 */
    float px, py, pz;
    float bound2;        /* 0.25 * dot(size, size) (shader.comp:887) */
    float m[9];          /* RotationMatrix(rotation), column-major: m[3*col + row] */
    float sx, sy, sz;    /* size */
    float materialID, lightID;
    float pad[2];
} PtDevBox;

typedef struct PtDevLens { /* shader.comp:90-99, 366-448, 895-910 */
    float px, py, pz;
    float bound2;        /* bounding-sphere radius^2 (shader.comp:899-905) */
    float m[9];          /* RotationMatrix(rotation) */
    float sradius;       /* slice.radius = 2 * focalLength */
    float sradius2;      /* sradius * sradius */
    float sliceOffset;   /* sradius - lensThicknessHalf */
    float shift;         /* localSlicePos - sliceSize - sliceOffset (shader.comp:377,379) */
    float invertSide;    /* 1 when !isConverging (isSideInvert), else 0 */
    float materialID, lightID;
} PtDevLens;

typedef struct PtDevCyclide { /* shader.comp:101-112, 633-679 */
    float px, py, pz;
    float brad;          /* packed brad^2 * maxscale^2 (host:3723-3725) */
    float m[9];
    float sx, sy, sz;    /* scale */
    float a, b, c, d;
    float materialID, lightID;
    float pad[2];
} PtDevCyclide;

typedef struct PtDevSdf { /* shader.comp:114-117: bounding box */
    float px, py, pz, sx, sy, sz, pad[2];
} PtDevSdf;

typedef struct PtDevLightSlot { /* what SampleRandomLightSource (shader.comp:1225-1285) returns for lightIDs[j] */
    float px, py, pz;
    float boundingRadius;
    float lightID;       /* lightIDOut */
    int objectID;        /* returned lightObjectID */
    float pad[2];
} PtDevLightSlot;

/* Surface extension of one material (mirror / glossy / dielectric surfaces -- NOT in the reference, whose
 * every surface is the Lambertian of shader.comp:1075-1091; pt_set_surface_ext, pt_abi.h).  Entry i extends material i
 * (0-based, the scene file's order = floor(materialID) in the kernel). */
#define PT_DEV_MAX_SURFACE_EXT 64
typedef struct PtDevSurfaceExt {
    int bsdf;            /* pt_bsdf: 0 the reference's Lambertian, 1 mirror, 2 glossy (GGX conductor), 3 dielectric */
    float roughness;     /* glossy: GGX alpha = roughness^2 */
    float ior;           /* dielectric: index of refraction; 0 = BK7 Sellmeier at the hero wavelength (shader.comp:1064-1073) */
    float pad;
} PtDevSurfaceExt;

typedef struct PtDevScene {
    int nSpheres, nPlanes, nBoxes, nLenses, nCyclides, nSdfs, nLightSlots;
    float numLights;     /* numObjects[6] as the float the shader multiplies with */
    float invNumLights;  /* 1.0 / numObjects[6] (shader.comp:1289) */
    /* offsets (in floats) of each type's records inside pool; spheres start at 0 */
    int offPlanes, offBoxes, offLenses, offCyclides, offSdfs;
    int nSurfaceExt;     /* 0: every surface is the reference's (the kernels are then built without the extension code) */
    int pad0;
    float pool[PT_DEV_POOL_FLOATS]; /* PtDevSphere[], PtDevPlane[], PtDevBox[], PtDevLens[], PtDevCyclide[], PtDevSdf[] */
    PtDevLightSlot lightSlots[PT_DEV_MAX_LIGHT_SLOTS];
    PtDevSurfaceExt surfaceExt[PT_DEV_MAX_SURFACE_EXT];
} PtDevScene;

/* Pointer types are spelled out on the device and opaque on the host (same layout). */
#if defined(__CUDACC__) || defined(__CUDACC_RTC__) || defined(PT_CUDA_SHIM_H) /* the device, or tests/simt */
#define PT_WF_PTR(T) T*
#else
#define PT_WF_PTR(T) void*
#endif

/* Per-dispatch constants: the push constants (shader.comp:31-52) plus what Scene()/TracePathLens() derive from them */
typedef struct PtDevParams {
    int width, height;
    int firstSample;       /* frame - samplesPerFrame: sample index of k = 0 (shader.comp:954) */
    int samplesPerFrame;
    int pathLength;
    int accumMode;         /* 0 static running mean, 1 temporal EMA (shader.comp:1500-1506), 2 raw sum */
    float accumN;          /* float(unitSamples) (mode 0, shader.comp:1504-1505) */
    float accumNm1;        /* float(unitSamples - 1) */
    float accumWeight;     /* EMA weight (mode 1) */
    float spfFloat;        /* float(samplesPerFrame) */
    float exposure;        /* apertureSize * apertureSize * float(ISO) (shader.comp:1519) */
    float resX, resY;      /* float(resolution) */
    float camPosX, camPosY, camPosZ;
    float sensorScale;     /* -cameraSize * 0.5 (shader.comp:1457) */
    float halfAperture;    /* 0.5 * apertureSize (shader.comp:1461) */
    float apertureDist;
    float camM[9];         /* RotationMatrix(vec3(cameraAngle, 0)), column-major */
    PtDevLens camLens;     /* TracePathLens' lens object (shader.comp:1411-1418) */
    /* Camera rays generated ahead of the megakernel (option "pregen", pt_kernel.cuh pt_gen_body): two planes of
     * genCount float4 each -- (origin.xyz, dir.x) and (dir.y, dir.z, hero wavelength, seed bits) -- one record per
     * (warp tile, sample of the dispatch, pixel of the tile): record = (tile * samplesPerFrame + k) * 32 + pixel. */
    PT_WF_PTR(float4) gen;
    PT_WF_PTR(float4) rad; /* option "resolve": genCount float4, the radiance bundle of every finished sample (same index) */
    unsigned long long genCount;
    int blockY0;           /* first CTA row (8 pixel rows each) of this launch: a dispatch whose records outgrow the scratch
                              buffer runs as several bands of CTA rows */
    int pad1[3];
} PtDevParams;

/* Wavefront pipeline (pt_wavefront.cuh): device buffers of the path state, SoA. */
typedef struct PtWf {
    PT_WF_PTR(float4) rayO; PT_WF_PTR(float4) rayD; PT_WF_PTR(float4) wl; PT_WF_PTR(float4) rad; PT_WF_PTR(float4) thr;
    PT_WF_PTR(float4) shD; PT_WF_PTR(float4) shC; PT_WF_PTR(float4) hit0; PT_WF_PTR(float4) hit1; PT_WF_PTR(float4) col;
    PT_WF_PTR(float4) acc;
    PT_WF_PTR(uint2) misc;
    PT_WF_PTR(unsigned) qA; PT_WF_PTR(unsigned) qB; PT_WF_PTR(unsigned) qS; PT_WF_PTR(unsigned) qM;
    PT_WF_PTR(unsigned) cnt;  /* [0] nA  [1] nB  [2] nS  [3] nM  [4] march head */
    unsigned P, nPix;        /* paths in flight (pixels x samples of the chunk), pixels of the chunk's band */
    int chunkBase;           /* k of the chunk's first sample (sample index = firstSample + k) */
    int chunkSamples;
    int lastChunk;
    unsigned pixBase;        /* first pixel (row-major index) of the chunk's band: chunks tile the frame so that the path
                                state of one chunk stays resident in L2 */
} PtWf;

/* raw tables the kernel indexes with computed ids; a copy of the tail of the uniform block in global memory:
 * flat float[4097] exactly like pt_ubo, so clamped flat indexing matches the oracle's Shader::at() */
#define PT_UBO_FLOATS 4097
#define PT_OFF_OBJ 7
#define PT_OFF_SDF (7 + 1024)
#define PT_OFF_MAT (PT_OFF_SDF + 768)
#define PT_OFF_LGT (PT_OFF_MAT + 783)
#define PT_OFF_LID (PT_OFF_LGT + 128)
#define PT_OFF_CIE (PT_OFF_LID + 64)

#endif /* PT_DEV_SCENE_H */
