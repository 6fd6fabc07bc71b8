/* binding.c -- CPython module `binding` for the B200 drone envs (host code in C).
 *
 * Mirrors the module the reference builds from pufferlib/ocean/env_binding.h +
 * drone_race/binding.c (or drone_swarm/binding.c): the same 15 method names
 * (EB:644-662), the same positional signatures, the same exception types for
 * the same mistakes (pinned by the reference's tests/test_env_binding.py), and
 * the same dict keys from vec_log.  The difference is what a handle stands
 * for: env structs live on the GPU behind libb200drone.so (include/b200drone.h)
 * and this file only parses Python arguments and forwards raw pointers.
 *
 * Buffers may be NumPy arrays (the reference's contract; stepped through
 * b2d_vec_step_host with H2D/D2H copies) or any object exposing
 * __cuda_array_interface__ (torch CUDA tensors, CuPy; zero-copy, asynchronous).
 *
 * Compile with -DB2D_BINDING_SWARM for the drone_swarm flavour.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#include <limits.h>
#include <stdint.h>
#include <string.h>

#include "../../include/b200drone.h"

#ifdef B2D_BINDING_SWARM
#define OBS_DIM B2D_SWARM_OBS
#else
#define OBS_DIM B2D_RACE_OBS
#endif

/* ---- what a Python-side handle points at ------------------------------------ */
typedef struct {
    void *obs, *act, *rew, *term, *trunc;
    int location; /* b2d_mem */
    int device;
    Py_ssize_t rows; /* rows of the slices handed to env_init */
} BufSet;

typedef struct VecH {
    uint32_t magic;
    b2d_vec *vec;
    void *stream;
    int num_envs;
    int rows_per_env;
    int location;
} VecH;

typedef struct EnvH {
    uint32_t magic;
    BufSet bufs;
    int seed;
    int max_rings, max_moves, num_agents;
    int device, env_id_base, math, write_clamped; /* optional kwargs (placement / arithmetic), kept for vectorize() */
    VecH *parent; /* set by vectorize(); or a private 1-env vec for env_reset/env_step */
    int index;
    int owns_parent;
} EnvH;

#define ENV_MAGIC 0xB2D0E001u
#define VEC_MAGIC 0xB2D0E002u

static int set_b2d_error(int rc) {
    const char *msg = b2d_last_error();
    if (rc == B2D_EINVAL) PyErr_SetString(PyExc_ValueError, msg);
    else if (rc == B2D_ENOMEM) PyErr_SetString(PyExc_MemoryError, msg);
    else PyErr_SetString(PyExc_RuntimeError, msg);
    return -1;
}

/* ---- argument parsing --------------------------------------------------------- */
/* One contract buffer: NumPy array or __cuda_array_interface__ object.
 * Error strings and exception types follow EB:62-135. */
static int parse_buffer(PyObject *o, const char *name, int want_1d, int reject_f64, int min_2d, void **ptr,
                        int *location, Py_ssize_t *rows, Py_ssize_t *stride0, Py_ssize_t *itemsize, Py_ssize_t *inner_out) {
    char msg[128];
    if (PyObject_TypeCheck(o, &PyArray_Type)) {
        PyArrayObject *a = (PyArrayObject *)o;
        if (!PyArray_ISCONTIGUOUS(a)) {
            snprintf(msg, sizeof msg, "%s must be contiguous", name);
            PyErr_SetString(PyExc_ValueError, msg);
            return -1;
        }
        if (reject_f64 && PyArray_ITEMSIZE(a) == sizeof(double)) {
            PyErr_SetString(PyExc_ValueError, "Action tensor passed as float64 (pass np.float32 buffer)");
            return -1;
        }
        if (want_1d && PyArray_NDIM(a) != 1) {
            snprintf(msg, sizeof msg, "%s must be 1D", name);
            PyErr_SetString(PyExc_ValueError, msg);
            return -1;
        }
        if (min_2d && PyArray_NDIM(a) < 2) {
            PyErr_SetString(PyExc_ValueError, "Batched Observations must be at least 2D");
            return -1;
        }
        *ptr = PyArray_DATA(a);
        *location = B2D_MEM_HOST;
        *rows = PyArray_NDIM(a) > 0 ? PyArray_DIM(a, 0) : 1;
        *stride0 = PyArray_NDIM(a) > 0 ? PyArray_STRIDE(a, 0) : 0;
        *itemsize = PyArray_ITEMSIZE(a);
        *inner_out = 1;
        for (int k = 1; k < PyArray_NDIM(a); k++) *inner_out *= PyArray_DIM(a, k);
        return 0;
    }
    PyObject *cai = PyObject_GetAttrString(o, "__cuda_array_interface__");
    if (!cai) {
        PyErr_Clear();
        snprintf(msg, sizeof msg, "%s must be a NumPy array", name);
        PyErr_SetString(PyExc_TypeError, msg);
        return -1;
    }
    int ok = -1;
    PyObject *data = PyDict_GetItemString(cai, "data");
    PyObject *shape = PyDict_GetItemString(cai, "shape");
    PyObject *strides = PyDict_GetItemString(cai, "strides");
    PyObject *typestr = PyDict_GetItemString(cai, "typestr");
    if (!data || !shape || !PyTuple_Check(data) || !PyTuple_Check(shape)) {
        PyErr_SetString(PyExc_TypeError, "malformed __cuda_array_interface__");
        goto done;
    }
    if (strides && strides != Py_None) {
        /* only C-contiguous layouts are described with explicit strides by some producers */
        Py_ssize_t nd = PyTuple_Size(shape), expect = 0;
        const char *ts = typestr ? PyUnicode_AsUTF8(typestr) : NULL;
        expect = ts ? atoi(ts + 2) : 4;
        for (Py_ssize_t k = nd - 1; k >= 0; k--) {
            if (PyLong_AsSsize_t(PyTuple_GetItem(strides, k)) != expect) {
                snprintf(msg, sizeof msg, "%s must be contiguous", name);
                PyErr_SetString(PyExc_ValueError, msg);
                goto done;
            }
            expect *= PyLong_AsSsize_t(PyTuple_GetItem(shape, k));
        }
    }
    if (reject_f64 && typestr) {
        const char *ts = PyUnicode_AsUTF8(typestr);
        if (ts && strlen(ts) >= 3 && ts[1] == 'f' && ts[2] == '8') {
            PyErr_SetString(PyExc_ValueError, "Action tensor passed as float64 (pass np.float32 buffer)");
            goto done;
        }
    }
    {
        Py_ssize_t nd = PyTuple_Size(shape);
        if (want_1d && nd != 1) {
            snprintf(msg, sizeof msg, "%s must be 1D", name);
            PyErr_SetString(PyExc_ValueError, msg);
            goto done;
        }
        if (min_2d && nd < 2) {
            PyErr_SetString(PyExc_ValueError, "Batched Observations must be at least 2D");
            goto done;
        }
        *rows = nd > 0 ? PyLong_AsSsize_t(PyTuple_GetItem(shape, 0)) : 1;
        Py_ssize_t inner = 1;
        for (Py_ssize_t k = 1; k < nd; k++) inner *= PyLong_AsSsize_t(PyTuple_GetItem(shape, k));
        const char *ts = typestr ? PyUnicode_AsUTF8(typestr) : NULL;
        *itemsize = ts ? atoi(ts + 2) : 4;
        *inner_out = inner;
        *stride0 = inner * *itemsize;
    }
    *ptr = PyLong_AsVoidPtr(PyTuple_GetItem(data, 0));
    *location = B2D_MEM_DEVICE;
    ok = 0;
done:
    Py_DECREF(cai);
    return ok;
}

typedef struct {
    void *ptr[5];
    Py_ssize_t rows[5], stride[5], itemsize[5], inner[5];
    int location;
} FiveBufs;

static const char *BUF_NAMES[5] = {"Observations", "Actions", "Rewards", "Terminals", "Truncations"};

static int parse_five(PyObject *args, int batched, FiveBufs *f) {
    int loc[5];
    for (int k = 0; k < 5; k++) {
        int want_1d = k >= 2;
        if (parse_buffer(PyTuple_GetItem(args, k), BUF_NAMES[k], want_1d, k == 1, batched && k == 0, &f->ptr[k], &loc[k],
                         &f->rows[k], &f->stride[k], &f->itemsize[k], &f->inner[k]) < 0)
            return -1;
    }
    /* The library DMA-writes rows * OBS_DIM * 4, rows * 4 and rows bytes into these arrays: element size and
     * row width must be what the kernels assume (the reference touches one env's slice at a time and can
     * afford to be looser, EB:62-135). */
    {
        static const Py_ssize_t want_item[5] = {4, 4, 4, 1, 1};
        const Py_ssize_t want_inner[5] = {OBS_DIM, B2D_ACT, 1, 1, 1};
        char msg[160];
        for (int k = 0; k < 5; k++) {
            if (f->itemsize[k] != want_item[k]) {
                snprintf(msg, sizeof msg, "%s must have %d-byte elements (%s), got %d-byte elements", BUF_NAMES[k], (int)want_item[k],
                         k < 3 ? "float32" : "bool / uint8", (int)f->itemsize[k]);
                PyErr_SetString(PyExc_ValueError, msg);
                return -1;
            }
            if (f->inner[k] != want_inner[k]) {
                snprintf(msg, sizeof msg, "%s must have %d values per row, got %d", BUF_NAMES[k], (int)want_inner[k], (int)f->inner[k]);
                PyErr_SetString(PyExc_ValueError, msg);
                return -1;
            }
        }
    }
    for (int k = 1; k < 5; k++) {
        if (loc[k] != loc[0]) {
            PyErr_SetString(PyExc_ValueError, "all buffers must live in the same memory space (all NumPy or all CUDA)");
            return -1;
        }
    }
    f->location = loc[0];
    return 0;
}

/* EB:615-641 */
static double unpack(PyObject *kwargs, const char *key) {
    PyObject *val = kwargs ? PyDict_GetItemString(kwargs, key) : NULL;
    char msg[128];
    if (val == NULL) {
        snprintf(msg, sizeof msg, "Missing required keyword argument '%s'", key);
        PyErr_SetString(PyExc_TypeError, msg);
        return 1;
    }
    if (PyLong_Check(val)) {
        long out = PyLong_AsLong(val);
        if (out > INT_MAX || out < INT_MIN) {
            snprintf(msg, sizeof msg, "Value %ld of integer argument %s is out of range", out, key);
            PyErr_SetString(PyExc_TypeError, msg);
            return 1;
        }
        return (double)out;
    }
    if (PyFloat_Check(val)) return PyFloat_AsDouble(val);
    snprintf(msg, sizeof msg, "Failed to unpack keyword %s as int", key);
    PyErr_SetString(PyExc_TypeError, msg);
    return 1;
}

static int opt_int(PyObject *kwargs, const char *key, int dflt) {
    PyObject *val = kwargs ? PyDict_GetItemString(kwargs, key) : NULL;
    if (!val) return dflt;
    if (PyUnicode_Check(val)) { /* math="strict" */
        const char *s = PyUnicode_AsUTF8(val);
        if (s && strcmp(s, "strict") == 0) return B2D_MATH_STRICT;
        if (s && strcmp(s, "fast") == 0) return B2D_MATH_FAST;
    }
    long v = PyLong_AsLong(val);
    if (PyErr_Occurred()) {
        PyErr_Clear();
        return dflt;
    }
    return (int)v;
}

/* my_init of DR/binding.c:6-11 / DS/binding.c:6-11: the required kwargs */
static int parse_env_kwargs(PyObject *kwargs, int *max_rings, int *max_moves, int *num_agents) {
#ifdef B2D_BINDING_SWARM
    *num_agents = (int)unpack(kwargs, "num_agents");
    if (PyErr_Occurred()) return -1;
    *max_rings = (int)unpack(kwargs, "max_rings");
    if (PyErr_Occurred()) return -1;
    *max_moves = 0;
#else
    *max_rings = (int)unpack(kwargs, "max_rings");
    if (PyErr_Occurred()) return -1;
    *max_moves = (int)unpack(kwargs, "max_moves");
    if (PyErr_Occurred()) return -1;
    *num_agents = 1;
#endif
    return 0;
}

typedef struct {
    int device, env_id_base, math, write_clamped;
} Extra;

/* placement / arithmetic kwargs that have no counterpart in the reference: device=0, env_id_base=0,
 * math="fast"|"strict", write_clamped_actions.  The reference clamps the shared action buffer in
 * place (DR/dronelib.h:437), so after a step the caller's `actions` array holds clamp(action, -1, 1).
 * write_clamped_actions=-1 (default) means "like the reference where the caller can see it": NumPy
 * buffers get the clamped values back (host-side clamp of the caller-visible array, no extra PCIe
 * traffic); device buffers are left untouched unless write_clamped_actions=1 (a policy's output
 * tensor is usually not the env's to modify, and the store costs 16 B/env of HBM traffic). */
static Extra parse_extra(PyObject *kwargs) {
    Extra x;
    x.device = opt_int(kwargs, "device", 0);
    x.env_id_base = opt_int(kwargs, "env_id_base", 0);
    x.math = opt_int(kwargs, "math", B2D_MATH_FAST);
    x.write_clamped = opt_int(kwargs, "write_clamped_actions", -1);
    return x;
}

static VecH *make_vec(const FiveBufs *f, int num_envs, int seed, int max_rings, int max_moves, int num_agents,
                      Extra x) {
    b2d_buffers ext;
    ext.observations = (float *)f->ptr[0];
    ext.actions = (float *)f->ptr[1];
    ext.rewards = (float *)f->ptr[2];
    ext.terminals = (unsigned char *)f->ptr[3];
    ext.truncations = (unsigned char *)f->ptr[4];
    ext.location = f->location;
    VecH *vh = (VecH *)calloc(1, sizeof(VecH));
    if (!vh) {
        PyErr_SetString(PyExc_MemoryError, "Failed to allocate vec env");
        return NULL;
    }
    int rc;
#ifdef B2D_BINDING_SWARM
    b2d_swarm_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.num_envs = num_envs;
    cfg.num_agents = num_agents;
    cfg.max_rings = max_rings;
    cfg.device = x.device;
    cfg.seed = (uint64_t)(int64_t)seed;
    cfg.env_id_base = (uint32_t)x.env_id_base;
    cfg.math = x.math;
    cfg.write_clamped_actions = x.write_clamped;
    (void)max_moves;
    Py_BEGIN_ALLOW_THREADS rc = b2d_swarm_create(&vh->vec, &cfg, &ext);
    Py_END_ALLOW_THREADS
#else
    b2d_race_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.num_envs = num_envs;
    cfg.max_rings = max_rings;
    cfg.max_moves = max_moves;
    cfg.device = x.device;
    cfg.seed = (uint64_t)(int64_t)seed;
    cfg.env_id_base = (uint32_t)x.env_id_base;
    cfg.math = x.math;
    cfg.write_clamped_actions = x.write_clamped;
    (void)num_agents;
    Py_BEGIN_ALLOW_THREADS rc = b2d_race_create(&vh->vec, &cfg, &ext);
    Py_END_ALLOW_THREADS
#endif
    if (rc != B2D_OK) {
        free(vh);
        set_b2d_error(rc);
        return NULL;
    }
    vh->magic = VEC_MAGIC;
    vh->num_envs = num_envs;
    vh->rows_per_env = num_agents;
    vh->location = f->location;
    vh->stream = NULL;
    return vh;
}

static EnvH *unpack_env(PyObject *args) {
    PyObject *h = PyTuple_GetItem(args, 0);
    if (!h || !PyObject_TypeCheck(h, &PyLong_Type)) {
        PyErr_SetString(PyExc_TypeError, "env_handle must be an integer");
        return NULL;
    }
    EnvH *env = (EnvH *)PyLong_AsVoidPtr(h);
    if (!env || env->magic != ENV_MAGIC) {
        PyErr_SetString(PyExc_ValueError, "Invalid env handle");
        return NULL;
    }
    return env;
}

static VecH *unpack_vecenv(PyObject *args) {
    PyObject *h = PyTuple_GetItem(args, 0);
    if (!h || !PyObject_TypeCheck(h, &PyLong_Type)) {
        PyErr_SetString(PyExc_TypeError, "env_handle must be an integer");
        return NULL;
    }
    VecH *vec = (VecH *)PyLong_AsVoidPtr(h);
    if (!vec || vec->magic != VEC_MAGIC || vec->num_envs <= 0) {
        PyErr_SetString(PyExc_ValueError, "Missing or invalid vec env handle");
        return NULL;
    }
    return vec;
}

/* ---- env_* ------------------------------------------------------------------------- */
/* EB:50-174.  No device work happens here: the slot is materialised by vectorize()
 * (or lazily as a vec of one by env_reset / env_step). */
static PyObject *env_init(PyObject *self, PyObject *args, PyObject *kwargs) {
    if (PyTuple_Size(args) != 6) {
        PyErr_SetString(PyExc_TypeError, "Environment requires 5 arguments");
        return NULL;
    }
    FiveBufs f;
    if (parse_five(args, 0, &f) < 0) return NULL;
    PyObject *seed_arg = PyTuple_GetItem(args, 5);
    if (!PyObject_TypeCheck(seed_arg, &PyLong_Type)) {
        PyErr_SetString(PyExc_TypeError, "seed must be an integer");
        return NULL;
    }
    EnvH *env = (EnvH *)calloc(1, sizeof(EnvH));
    if (!env) {
        PyErr_SetString(PyExc_MemoryError, "Failed to allocate environment");
        return NULL;
    }
    env->seed = (int)PyLong_AsLong(seed_arg);
    if (parse_env_kwargs(kwargs, &env->max_rings, &env->max_moves, &env->num_agents) < 0) {
        free(env);
        return NULL;
    }
    env->magic = ENV_MAGIC;
    env->bufs.obs = f.ptr[0]; env->bufs.act = f.ptr[1]; env->bufs.rew = f.ptr[2];
    env->bufs.term = f.ptr[3]; env->bufs.trunc = f.ptr[4];
    env->bufs.location = f.location;
    {
        Extra x = parse_extra(kwargs);
        env->device = x.device; env->env_id_base = x.env_id_base; env->math = x.math; env->write_clamped = x.write_clamped;
    }
    env->bufs.device = env->device;
    env->bufs.rows = f.rows[0];
    return PyLong_FromVoidPtr(env);
}

static int env_materialise(EnvH *env) {
    if (env->parent) return 0;
    FiveBufs f;
    f.ptr[0] = env->bufs.obs; f.ptr[1] = env->bufs.act; f.ptr[2] = env->bufs.rew;
    f.ptr[3] = env->bufs.term; f.ptr[4] = env->bufs.trunc;
    f.location = env->bufs.location;
    Extra x = {env->device, env->env_id_base, env->math, env->write_clamped};
    VecH *vh = make_vec(&f, 1, env->seed, env->max_rings, env->max_moves, env->num_agents, x);
    if (!vh) return -1;
    env->parent = vh;
    env->index = 0;
    env->owns_parent = 1;
    return 0;
}

static int vec_reset_impl(VecH *vh, uint64_t seed) {
    int rc;
    Py_BEGIN_ALLOW_THREADS
    if (vh->location == B2D_MEM_HOST) rc = b2d_vec_reset_host(vh->vec, seed, vh->stream);
    else rc = b2d_vec_reset(vh->vec, seed, vh->stream);
    Py_END_ALLOW_THREADS
    return rc == B2D_OK ? 0 : set_b2d_error(rc);
}

static int vec_step_impl(VecH *vh) {
    int rc;
    if (vh->location == B2D_MEM_HOST) {
        Py_BEGIN_ALLOW_THREADS rc = b2d_vec_step_host(vh->vec, vh->stream);
        Py_END_ALLOW_THREADS
    } else {
        rc = b2d_vec_step(vh->vec, vh->stream); /* async launch: microseconds, keep the GIL */
    }
    return rc == B2D_OK ? 0 : set_b2d_error(rc);
}

/* EB:177-189 */
static PyObject *env_reset(PyObject *self, PyObject *args) {
    if (PyTuple_Size(args) != 2) {
        PyErr_SetString(PyExc_TypeError, "env_reset requires 2 arguments");
        return NULL;
    }
    EnvH *env = unpack_env(args);
    if (!env) return NULL;
    if (env->parent && !env->owns_parent) {
        PyErr_SetString(PyExc_ValueError, "env belongs to a vectorized handle: use vec_reset");
        return NULL;
    }
    PyObject *seed_arg = PyTuple_GetItem(args, 1);
    uint64_t seed = PyLong_Check(seed_arg) ? (uint64_t)PyLong_AsLongLong(seed_arg) : (uint64_t)env->seed;
    if (env_materialise(env) < 0 || vec_reset_impl(env->parent, seed) < 0) return NULL;
    Py_RETURN_NONE;
}

/* EB:192-205 */
static PyObject *env_step(PyObject *self, PyObject *args) {
    if (PyTuple_Size(args) != 1) {
        PyErr_SetString(PyExc_TypeError, "vec_render requires 1 argument");
        return NULL;
    }
    EnvH *env = unpack_env(args);
    if (!env) return NULL;
    if (env->parent && !env->owns_parent) {
        PyErr_SetString(PyExc_ValueError, "env belongs to a vectorized handle: use vec_step");
        return NULL;
    }
    if (env_materialise(env) < 0 || vec_step_impl(env->parent) < 0) return NULL;
    Py_RETURN_NONE;
}

/* EB:208-215; rendering (raylib) is outside this library */
static PyObject *env_render(PyObject *self, PyObject *args) {
    if (!unpack_env(args)) return NULL;
    Py_RETURN_NONE;
}

/* EB:218-226 */
static PyObject *env_close(PyObject *self, PyObject *args) {
    EnvH *env = unpack_env(args);
    if (!env) return NULL;
    if (env->parent && env->owns_parent) {
        b2d_vec_close(env->parent->vec);
        env->parent->magic = 0;
        free(env->parent);
    }
    env->magic = 0;
    free(env);
    Py_RETURN_NONE;
}

/* EB:228-239: the reference's drone envs return {}; ours returns the state blob of
 * include/b200drone.h:b2d_get_state under "state" (parity hook / env checkpoint). */
static PyObject *env_get(PyObject *self, PyObject *args) {
    EnvH *env = unpack_env(args);
    if (!env) return NULL;
    PyObject *dict = PyDict_New();
    if (!env->parent) return dict;
    int nf = b2d_state_blob_floats(env->parent->vec);
    npy_intp dims[1] = {nf};
    PyObject *arr = PyArray_ZEROS(1, dims, NPY_FLOAT32, 0);
    if (!arr) return NULL;
    int id = env->index;
    int rc = b2d_get_state(env->parent->vec, &id, 1, (float *)PyArray_DATA((PyArrayObject *)arr));
    if (rc != B2D_OK) {
        Py_DECREF(arr);
        Py_DECREF(dict);
        set_b2d_error(rc);
        return NULL;
    }
    PyDict_SetItemString(dict, "state", arr);
    Py_DECREF(arr);
    return dict;
}

/* EB:241-260: env_put(handle, state=blob) */
static PyObject *env_put(PyObject *self, PyObject *args, PyObject *kwargs) {
    if (PyTuple_Size(args) != 1) {
        PyErr_SetString(PyExc_TypeError, "env_put requires 1 positional argument");
        return NULL;
    }
    EnvH *env = unpack_env(args);
    if (!env) return NULL;
    PyObject *st = kwargs ? PyDict_GetItemString(kwargs, "state") : NULL;
    if (!st) Py_RETURN_NONE;
    if (env_materialise(env) < 0) return NULL;
    PyArrayObject *arr = (PyArrayObject *)PyArray_FROM_OTF(st, NPY_FLOAT32, NPY_ARRAY_IN_ARRAY);
    if (!arr) return NULL;
    int nf = b2d_state_blob_floats(env->parent->vec);
    if (PyArray_SIZE(arr) != nf) {
        Py_DECREF(arr);
        PyErr_SetString(PyExc_ValueError, "state blob has the wrong length");
        return NULL;
    }
    int id = env->index;
    int rc = b2d_put_state(env->parent->vec, &id, 1, (const float *)PyArray_DATA(arr));
    Py_DECREF(arr);
    if (rc != B2D_OK) {
        set_b2d_error(rc);
        return NULL;
    }
    Py_RETURN_NONE;
}

/* ---- vec_* ------------------------------------------------------------------------- */
/* EB:288-446 */
static PyObject *vec_init(PyObject *self, PyObject *args, PyObject *kwargs) {
    if (PyTuple_Size(args) != 7) {
        PyErr_SetString(PyExc_TypeError, "vec_init requires 6 arguments");
        return NULL;
    }
    PyObject *num_envs_arg = PyTuple_GetItem(args, 5);
    if (!PyObject_TypeCheck(num_envs_arg, &PyLong_Type)) {
        PyErr_SetString(PyExc_TypeError, "num_envs must be an integer");
        return NULL;
    }
    int num_envs = (int)PyLong_AsLong(num_envs_arg);
    if (num_envs <= 0) {
        PyErr_SetString(PyExc_TypeError, "num_envs must be greater than 0");
        return NULL;
    }
    PyObject *seed_obj = PyTuple_GetItem(args, 6);
    if (!PyObject_TypeCheck(seed_obj, &PyLong_Type)) {
        PyErr_SetString(PyExc_TypeError, "seed must be an integer");
        return NULL;
    }
    int seed = (int)PyLong_AsLong(seed_obj);
    FiveBufs f;
    if (parse_five(args, 1, &f) < 0) return NULL;
    int max_rings, max_moves, num_agents;
    if (parse_env_kwargs(kwargs, &max_rings, &max_moves, &num_agents) < 0) return NULL;
    for (int k = 0; k < 5; k++) {
        if (f.rows[k] < (Py_ssize_t)num_envs * num_agents) {
            PyErr_SetString(PyExc_ValueError, "buffers are smaller than num_envs rows");
            return NULL;
        }
    }
    VecH *vh = make_vec(&f, num_envs, seed, max_rings, max_moves, num_agents, parse_extra(kwargs));
    if (!vh) return NULL;
    return PyLong_FromVoidPtr(vh);
}

/* EB:450-480.  The handles must describe consecutive slices of one set of buffers (what
 * DR/drone_race.py:37-51 passes); they become ONE device-resident vector. */
static PyObject *vectorize(PyObject *self, PyObject *args) {
    Py_ssize_t num_envs = PyTuple_Size(args);
    if (num_envs == 0) {
        PyErr_SetString(PyExc_TypeError, "make_vec requires at least 1 env id");
        return NULL;
    }
    EnvH **envs = (EnvH **)calloc((size_t)num_envs, sizeof(EnvH *));
    if (!envs) {
        PyErr_SetString(PyExc_MemoryError, "Failed to allocate vec env");
        return NULL;
    }
    for (Py_ssize_t i = 0; i < num_envs; i++) {
        PyObject *h = PyTuple_GetItem(args, i);
        if (!PyObject_TypeCheck(h, &PyLong_Type)) {
            free(envs);
            PyErr_SetString(PyExc_TypeError, "Env ids must be integers. Pass them as separate args with *env_ids, not as a list.");
            return NULL;
        }
        envs[i] = (EnvH *)PyLong_AsVoidPtr(h);
        if (!envs[i] || envs[i]->magic != ENV_MAGIC || envs[i]->parent) {
            free(envs);
            PyErr_SetString(PyExc_ValueError, "Invalid env handle");
            return NULL;
        }
    }
    EnvH *e0 = envs[0];
    const Py_ssize_t rpe = e0->num_agents;
    for (Py_ssize_t i = 1; i < num_envs; i++) {
        EnvH *e = envs[i];
        int same = e->max_rings == e0->max_rings && e->max_moves == e0->max_moves && e->num_agents == e0->num_agents &&
                   e->bufs.location == e0->bufs.location &&
                   (char *)e->bufs.obs == (char *)e0->bufs.obs + i * rpe * OBS_DIM * 4 &&
                   (char *)e->bufs.act == (char *)e0->bufs.act + i * rpe * 16 &&
                   (char *)e->bufs.rew == (char *)e0->bufs.rew + i * rpe * 4 &&
                   (char *)e->bufs.term == (char *)e0->bufs.term + i * rpe &&
                   (char *)e->bufs.trunc == (char *)e0->bufs.trunc + i * rpe;
        if (!same) {
            free(envs);
            PyErr_SetString(PyExc_ValueError, "vectorize needs envs built on consecutive slices of the same buffers with identical kwargs");
            return NULL;
        }
    }
    FiveBufs f;
    f.ptr[0] = e0->bufs.obs; f.ptr[1] = e0->bufs.act; f.ptr[2] = e0->bufs.rew;
    f.ptr[3] = e0->bufs.term; f.ptr[4] = e0->bufs.trunc;
    f.location = e0->bufs.location;
    Extra x0 = {e0->device, e0->env_id_base, e0->math, e0->write_clamped};
    VecH *vh = make_vec(&f, (int)num_envs, e0->seed, e0->max_rings, e0->max_moves, e0->num_agents, x0);
    if (!vh) {
        free(envs);
        return NULL;
    }
    for (Py_ssize_t i = 0; i < num_envs; i++) {
        envs[i]->parent = vh;
        envs[i]->index = (int)i;
        envs[i]->owns_parent = 0;
    }
    free(envs);
    return PyLong_FromVoidPtr(vh);
}

/* EB:482-506 */
static PyObject *vec_reset(PyObject *self, PyObject *args) {
    if (PyTuple_Size(args) != 2) {
        PyErr_SetString(PyExc_TypeError, "vec_reset requires 2 arguments");
        return NULL;
    }
    VecH *vh = unpack_vecenv(args);
    if (!vh) return NULL;
    PyObject *seed_arg = PyTuple_GetItem(args, 1);
    if (!PyObject_TypeCheck(seed_arg, &PyLong_Type)) {
        PyErr_SetString(PyExc_TypeError, "seed must be an integer");
        return NULL;
    }
    if (vec_reset_impl(vh, (uint64_t)PyLong_AsLongLong(seed_arg)) < 0) return NULL;
    Py_RETURN_NONE;
}

/* EB:508-524 */
static PyObject *vec_step(PyObject *self, PyObject *args) {
    if (PyTuple_Size(args) != 1) {
        PyErr_SetString(PyExc_TypeError, "vec_step requires 1 argument");
        return NULL;
    }
    VecH *vh = unpack_vecenv(args);
    if (!vh) return NULL;
    if (vec_step_impl(vh) < 0) return NULL;
    Py_RETURN_NONE;
}

/* extension: vec_step_actions(vec, actions) == `buffer_actions[:] = actions; vec_step(vec)` of the
 * reference wrappers (DR/drone_race.py:59-62) in one call for NumPy-buffer handles: the copy into
 * the shared action buffer runs on several cores with the GIL released. */
static PyObject *vec_step_actions(PyObject *self, PyObject *args) {
    if (PyTuple_Size(args) != 2) {
        PyErr_SetString(PyExc_TypeError, "vec_step_actions requires 2 arguments");
        return NULL;
    }
    VecH *vh = unpack_vecenv(args);
    if (!vh) return NULL;
    PyObject *o = PyTuple_GetItem(args, 1);
    if (vh->location != B2D_MEM_HOST || !PyObject_TypeCheck(o, &PyArray_Type)) {
        PyErr_SetString(PyExc_TypeError, "vec_step_actions needs a NumPy-buffer handle and a NumPy array");
        return NULL;
    }
    PyArrayObject *a = (PyArrayObject *)o;
    if (!PyArray_ISCONTIGUOUS(a) || PyArray_TYPE(a) != NPY_FLOAT32 ||
        PyArray_SIZE(a) != (npy_intp)b2d_num_agents(vh->vec) * 4) {
        PyErr_SetString(PyExc_ValueError, "actions must be a contiguous float32 array of shape [num_agents, 4]");
        return NULL;
    }
    int rc;
    const float *src = (const float *)PyArray_DATA(a);
    Py_BEGIN_ALLOW_THREADS rc = b2d_vec_step_host_from(vh->vec, src, vh->stream);
    Py_END_ALLOW_THREADS
    if (rc != B2D_OK) {
        set_b2d_error(rc);
        return NULL;
    }
    Py_RETURN_NONE;
}

/* EB:526-548 */
static PyObject *vec_render(PyObject *self, PyObject *args) {
    if (PyTuple_Size(args) != 2) {
        PyErr_SetString(PyExc_TypeError, "vec_render requires 2 arguments");
        return NULL;
    }
    if (!unpack_vecenv(args)) return NULL;
    if (!PyObject_TypeCheck(PyTuple_GetItem(args, 1), &PyLong_Type)) {
        PyErr_SetString(PyExc_TypeError, "env_id must be an integer");
        return NULL;
    }
    Py_RETURN_NONE;
}

static int assign_to_dict(PyObject *dict, const char *key, float value) {
    PyObject *v = PyFloat_FromDouble(value);
    if (!v) return 1;
    int rc = PyDict_SetItemString(dict, key, v);
    Py_DECREF(v);
    return rc < 0;
}

/* EB:564-598 + my_log (DR/binding.c:13-23, DS/binding.c:13-23) */
static PyObject *vec_log(PyObject *self, PyObject *args) {
    VecH *vh = unpack_vecenv(args);
    if (!vh) return NULL;
    float l[B2D_LOG_FIELDS];
    int rc;
    Py_BEGIN_ALLOW_THREADS rc = b2d_vec_log(vh->vec, l, vh->stream);
    Py_END_ALLOW_THREADS
    if (rc != B2D_OK) {
        set_b2d_error(rc);
        return NULL;
    }
    PyObject *dict = PyDict_New();
    if (l[8] == 0.0f) return dict;
    assign_to_dict(dict, "perf", l[7]);
    assign_to_dict(dict, "score", l[6]);
#ifdef B2D_BINDING_SWARM
    assign_to_dict(dict, "rings_passed", l[2]);
    assign_to_dict(dict, "collision_rate", l[3]);
    assign_to_dict(dict, "oob", l[4]);
#else
    assign_to_dict(dict, "collision_rate", l[3]);
    assign_to_dict(dict, "oob", l[4]);
    assign_to_dict(dict, "timeout", l[5]);
#endif
    assign_to_dict(dict, "episode_return", l[0]);
    assign_to_dict(dict, "episode_length", l[1]);
    assign_to_dict(dict, "n", l[8]);
    return dict;
}

/* EB:600-613 */
static PyObject *vec_close(PyObject *self, PyObject *args) {
    VecH *vh = unpack_vecenv(args);
    if (!vh) return NULL;
    int rc;
    Py_BEGIN_ALLOW_THREADS rc = b2d_vec_close(vh->vec);
    Py_END_ALLOW_THREADS
    vh->magic = 0;
    free(vh);
    if (rc != B2D_OK) {
        set_b2d_error(rc);
        return NULL;
    }
    Py_RETURN_NONE;
}

/* EB:8-13: no shared state for the drone envs */
static PyObject *my_shared(PyObject *self, PyObject *args, PyObject *kwargs) { Py_RETURN_NONE; }

/* ---- extensions (the reference's MY_METHODS slot, EB:29-31) -------------------------- */
/* vec_handle(vec) -> int: the raw b2d_vec* for ctypes / C callers (include/b200drone.h) */
static PyObject *vec_handle(PyObject *self, PyObject *args) {
    VecH *vh = unpack_vecenv(args);
    if (!vh) return NULL;
    return PyLong_FromVoidPtr(vh->vec);
}

/* vec_set_stream(vec, cuda_stream_ptr): stream for later vec_reset / vec_step / vec_log
 * (default 0 = the legacy default stream, which is also torch's default) */
static PyObject *vec_set_stream(PyObject *self, PyObject *args) {
    if (PyTuple_Size(args) != 2) {
        PyErr_SetString(PyExc_TypeError, "vec_set_stream requires 2 arguments");
        return NULL;
    }
    VecH *vh = unpack_vecenv(args);
    if (!vh) return NULL;
    vh->stream = PyLong_AsVoidPtr(PyTuple_GetItem(args, 1));
    if (PyErr_Occurred()) return NULL;
    Py_RETURN_NONE;
}

/* vec_buffers(vec) -> dict of raw device pointers of the contract buffers */
static PyObject *vec_buffers(PyObject *self, PyObject *args) {
    VecH *vh = unpack_vecenv(args);
    if (!vh) return NULL;
    b2d_buffers b;
    int rc = b2d_get_buffers(vh->vec, &b);
    if (rc != B2D_OK) {
        set_b2d_error(rc);
        return NULL;
    }
    PyObject *d = PyDict_New();
    PyObject *v;
#define PUT(k, p) v = PyLong_FromVoidPtr((void *)(p)); PyDict_SetItemString(d, k, v); Py_DECREF(v)
    PUT("observations", b.observations);
    PUT("actions", b.actions);
    PUT("rewards", b.rewards);
    PUT("terminals", b.terminals);
    PUT("truncations", b.truncations);
#undef PUT
    v = PyLong_FromLong(b2d_num_agents(vh->vec));
    PyDict_SetItemString(d, "num_agents", v);
    Py_DECREF(v);
    v = PyLong_FromLong(b2d_obs_dim(vh->vec));
    PyDict_SetItemString(d, "obs_dim", v);
    Py_DECREF(v);
    return d;
}

static PyMethodDef methods[] = {
    {"env_init", (PyCFunction)env_init, METH_VARARGS | METH_KEYWORDS, "Init environment with observation, action, reward, terminal, truncation arrays"},
    {"env_reset", env_reset, METH_VARARGS, "Reset the environment"},
    {"env_step", env_step, METH_VARARGS, "Step the environment"},
    {"env_render", env_render, METH_VARARGS, "Render the environment"},
    {"env_close", env_close, METH_VARARGS, "Close the environment"},
    {"env_get", env_get, METH_VARARGS, "Get the environment state"},
    {"env_put", (PyCFunction)env_put, METH_VARARGS | METH_KEYWORDS, "Put stuff into env"},
    {"vectorize", vectorize, METH_VARARGS, "Make a vector of environment handles"},
    {"vec_init", (PyCFunction)vec_init, METH_VARARGS | METH_KEYWORDS, "Initialize a vector of environments"},
    {"vec_reset", vec_reset, METH_VARARGS, "Reset the vector of environments"},
    {"vec_step", vec_step, METH_VARARGS, "Step the vector of environments"},
    {"vec_log", vec_log, METH_VARARGS, "Log the vector of environments"},
    {"vec_render", vec_render, METH_VARARGS, "Render the vector of environments"},
    {"vec_close", vec_close, METH_VARARGS, "Close the vector of environments"},
    {"shared", (PyCFunction)my_shared, METH_VARARGS | METH_KEYWORDS, "Shared state"},
    {"vec_step_actions", vec_step_actions, METH_VARARGS, "Copy actions into the shared buffer and step"},
    {"vec_handle", vec_handle, METH_VARARGS, "Raw b2d_vec* of a vector handle"},
    {"vec_set_stream", vec_set_stream, METH_VARARGS, "Set the CUDA stream of a vector handle"},
    {"vec_buffers", vec_buffers, METH_VARARGS, "Device pointers of the contract buffers"},
    {NULL, NULL, 0, NULL}};

static PyModuleDef module = {PyModuleDef_HEAD_INIT, "binding", NULL, -1, methods};

PyMODINIT_FUNC PyInit_binding(void) {
    import_array();
    return PyModule_Create(&module);
}
