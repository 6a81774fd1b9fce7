/* tabfast.c -- native core of yalps_b200.tableau.tableau_model (CPython extension `yalps_b200._tabfast`).
 *
 * Host-side producer at the drop-in boundary: turns the constraints / variables of a model into the ordered list of
 * stores `update(tableau, row, col, value)` that tableauModel makes into its zero-filled matrix (src/tableau.ts:73-134),
 * as (cell = row*width + col, value) pairs -- the form yalps_solve_sparse takes.  Same contract as the Python loops in
 * tableau.py (which stay as the fallback and as the specification the tests compare against): constraint keys merged in
 * first-seen order (:73-80), upper row before lower row (:82-86), coefficient stores in variable order with the
 * objective row first (:100-117), RHS stores (:119-127), `x <= 1` rows of the binaries (:130-134).  Keys are compared
 * the way a Python dict compares them.  Everything a model may legally contain but this file does not want to know
 * about (JS property order of integer-like dict keys, arbitrary iterables) is delegated to the `entries` callable.
 *
 * Build: see Makefile (gcc, Python.h); no CUDA here.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t *cell;
  double *val;
  Py_ssize_t n, cap;
} Stores;

/* cells must stay below 2^31: the reference indexes with Math.imul (src/tableau.ts:17) */
static int stores_push(Stores *s, int64_t cell, double v) {
  if (cell > INT32_MAX) {
    PyErr_SetString(PyExc_OverflowError, "tableau cell index does not fit int32 (height*width must be < 2^31)");
    return -1;
  }
  if (s->n == s->cap) {
    const Py_ssize_t cap = s->cap ? s->cap * 2 : 4096;
    int32_t *c = (int32_t *)realloc(s->cell, (size_t)cap * sizeof(int32_t));
    if (!c) {
      PyErr_NoMemory();
      return -1;
    }
    s->cell = c;
    double *v2 = (double *)realloc(s->val, (size_t)cap * sizeof(double));
    if (!v2) {
      PyErr_NoMemory();
      return -1;
    }
    s->val = v2;
    s->cap = cap;
  }
  s->cell[s->n] = (int32_t)cell;
  s->val[s->n] = v;
  s->n++;
  return 0;
}

/* float(obj) */
static int as_double(PyObject *o, double *out) {
  if (PyFloat_CheckExact(o)) {
    *out = PyFloat_AS_DOUBLE(o);
    return 0;
  }
  PyObject *f = PyNumber_Float(o);
  if (!f) return -1;
  *out = PyFloat_AS_DOUBLE(f);
  Py_DECREF(f);
  return 0;
}

/* constraint.name for a dict (missing -> None) or any object with attributes (missing -> None); new reference */
static PyObject *field(PyObject *con, PyObject *name) {
  if (PyDict_Check(con)) {
    PyObject *v = PyDict_GetItemWithError(con, name);
    if (!v) {
      if (PyErr_Occurred()) return NULL;
      Py_RETURN_NONE;
    }
    Py_INCREF(v);
    return v;
  }
  PyObject *v = PyObject_GetAttr(con, name);
  if (!v) {
    if (!PyErr_ExceptionMatches(PyExc_AttributeError)) return NULL;
    PyErr_Clear();
    Py_RETURN_NONE;
  }
  return v;
}

/* the two members of a (key, value) pair; borrowed references kept alive by *hold (released by the caller) */
static int unpack_pair(PyObject *item, PyObject **k, PyObject **v, PyObject **hold) {
  *hold = NULL;
  if (PyTuple_CheckExact(item) && PyTuple_GET_SIZE(item) == 2) {
    *k = PyTuple_GET_ITEM(item, 0);
    *v = PyTuple_GET_ITEM(item, 1);
    return 0;
  }
  if (PyList_CheckExact(item) && PyList_GET_SIZE(item) == 2) {
    *k = PyList_GET_ITEM(item, 0);
    *v = PyList_GET_ITEM(item, 1);
    return 0;
  }
  PyObject *seq = PySequence_Tuple(item); /* any iterable of exactly two, like `for key, value in ...` */
  if (!seq) return -1;
  if (PyTuple_GET_SIZE(seq) != 2) {
    Py_DECREF(seq);
    PyErr_SetString(PyExc_ValueError, "expected (key, value) pairs");
    return -1;
  }
  *k = PyTuple_GET_ITEM(seq, 0);
  *v = PyTuple_GET_ITEM(seq, 1);
  *hold = seq;
  return 0;
}

typedef struct {
  double *lower, *upper;
  int64_t *u_off, *l_off; /* row*width of the upper / lower row, -1 = none */
  char *is_obj;
  Py_ssize_t n, cap;
} Rows;

static int rows_grow(Rows *r) {
  if (r->n < r->cap) return 0;
  const Py_ssize_t cap = r->cap ? r->cap * 2 : 1024;
  void *p;
#define GROW(field, type)                                   \
  p = realloc(r->field, (size_t)cap * sizeof(type));        \
  if (!p) {                                                 \
    PyErr_NoMemory();                                       \
    return -1;                                              \
  }                                                         \
  r->field = (type *)p;
  GROW(lower, double)
  GROW(upper, double)
  GROW(u_off, int64_t)
  GROW(l_off, int64_t)
  GROW(is_obj, char)
#undef GROW
  r->cap = cap;
  return 0;
}

static int store_coef(Stores *st, const Rows *R, Py_ssize_t k, int64_t col, double sign, PyObject *coef_obj) {
  double coef;
  if (as_double(coef_obj, &coef)) return -1;
  if (R->is_obj[k] && stores_push(st, col, sign * coef)) return -1;
  if (R->u_off[k] >= 0 && stores_push(st, R->u_off[k] + col, coef)) return -1;
  if (R->l_off[k] >= 0 && stores_push(st, R->l_off[k] + col, -coef)) return -1;
  return 0;
}

/* one (constraint key, coefficient) pair of the variable in column col */
static int one_coef(Stores *st, const Rows *R, PyObject *index, int64_t col, double sign, PyObject *ckey, PyObject *coef) {
  PyObject *hit = PyDict_GetItemWithError(index, ckey); /* borrowed */
  if (!hit) return PyErr_Occurred() ? -1 : 0;           /* unknown key: ignored before the coefficient is looked at */
  const Py_ssize_t k = PyLong_AsSsize_t(hit);
  if (!R->is_obj[k] && R->u_off[k] < 0 && R->l_off[k] < 0) return 0; /* a key without a finite bound has no row either */
  return store_coef(st, R, k, col, sign, coef);
}

static int pairs_from_iterable(Stores *st, const Rows *R, PyObject *index, int64_t col, double sign, PyObject *iterable) {
  PyObject *it = PyObject_GetIter(iterable);
  if (!it) return -1;
  PyObject *item;
  int rc = 0;
  while (!rc && (item = PyIter_Next(it))) {
    PyObject *k, *v, *hold;
    rc = unpack_pair(item, &k, &v, &hold);
    if (!rc) rc = one_coef(st, R, index, col, sign, k, v);
    Py_XDECREF(hold);
    Py_DECREF(item);
  }
  Py_DECREF(it);
  if (!rc && PyErr_Occurred()) rc = -1;
  return rc;
}

/* a str key that starts with a digit may be a JS array index (ordered first, numerically): tableau.entries decides */
static int dict_needs_js_order(PyObject *d) {
  Py_ssize_t pos = 0;
  PyObject *k, *v;
  while (PyDict_Next(d, &pos, &k, &v)) {
    if (PyUnicode_Check(k) && PyUnicode_GET_LENGTH(k) > 0) {
      const Py_UCS4 c = PyUnicode_READ_CHAR(k, 0);
      if (c >= '0' && c <= '9') return 1;
    }
  }
  return 0;
}

static PyObject *s_equal, *s_min, *s_max;

/* build(variables, constraints, objective, sign, width, binary_cols, entries)
 *   variables    list of (key, coefficients)         constraints  list of (key, constraint), both already in entries() order
 *   -> (cells: bytes of int32, values: bytes of float64, rows: int = 1 + constraint rows) */
static PyObject *tab_build(PyObject *self, PyObject *args) {
  (void)self;
  PyObject *variables, *constraints, *objective, *binary_cols, *entries;
  double sign;
  long long width;
  if (!PyArg_ParseTuple(args, "O!O!OdLO!O", &PyList_Type, &variables, &PyList_Type, &constraints, &objective, &sign, &width,
                        &PyList_Type, &binary_cols, &entries))
    return NULL;
  Stores st = {0};
  Rows R = {0};
  PyObject *index = PyDict_New(), *result = NULL;
  if (!index) return NULL;

  /* ---- merge constraints per key, first-seen order (src/tableau.ts:73-80) */
  const Py_ssize_t ncons = PyList_GET_SIZE(constraints);
  for (Py_ssize_t i = 0; i < ncons; i++) {
    PyObject *key, *con, *hold;
    if (unpack_pair(PyList_GET_ITEM(constraints, i), &key, &con, &hold)) goto done;
    double lo = -INFINITY, hi = INFINITY;
    PyObject *eq = field(con, s_equal);
    int bad = !eq;
    if (!bad && eq != Py_None) {
      bad = as_double(eq, &lo);
      hi = lo;
    } else if (!bad) {
      PyObject *mn = field(con, s_min), *mx = mn ? field(con, s_max) : NULL;
      bad = !mn || !mx;
      if (!bad && mn != Py_None) bad = as_double(mn, &lo);
      if (!bad && mx != Py_None) bad = as_double(mx, &hi);
      Py_XDECREF(mn);
      Py_XDECREF(mx);
    }
    Py_XDECREF(eq);
    if (!bad) {
      PyObject *seen = PyDict_GetItemWithError(index, key);
      if (seen) {
        const Py_ssize_t k = PyLong_AsSsize_t(seen);
        if (lo > R.lower[k]) R.lower[k] = lo; /* max(lower, lo): Python's max keeps the first unless the second is greater */
        if (hi < R.upper[k]) R.upper[k] = hi;
      } else if (PyErr_Occurred()) {
        bad = 1;
      } else if (!(bad = rows_grow(&R))) {
        const Py_ssize_t k = R.n++;
        R.lower[k] = lo > -INFINITY ? lo : -INFINITY;
        R.upper[k] = hi < INFINITY ? hi : INFINITY;
        R.u_off[k] = R.l_off[k] = -1;
        R.is_obj[k] = 0;
        PyObject *num = PyLong_FromSsize_t(k);
        bad = !num || PyDict_SetItem(index, key, num);
        Py_XDECREF(num);
      }
    }
    Py_XDECREF(hold);
    if (bad) goto done;
  }

  /* ---- row numbering: upper row first, then lower row (src/tableau.ts:82-86) */
  int64_t rows = 1;
  const Py_ssize_t nkeys = R.n;
  for (Py_ssize_t k = 0; k < nkeys; k++) {
    if (isfinite(R.upper[k])) R.u_off[k] = rows++ * width;
    if (isfinite(R.lower[k])) R.l_off[k] = rows++ * width;
  }
  if (objective != Py_None) { /* the objective key may be a constraint key as well: both get their stores (:102-115) */
    PyObject *seen = PyDict_GetItemWithError(index, objective);
    if (seen) {
      R.is_obj[PyLong_AsSsize_t(seen)] = 1;
    } else {
      if (PyErr_Occurred() || rows_grow(&R)) goto done;
      const Py_ssize_t k = R.n++;
      R.lower[k] = -INFINITY;
      R.upper[k] = INFINITY;
      R.u_off[k] = R.l_off[k] = -1;
      R.is_obj[k] = 1;
      PyObject *num = PyLong_FromSsize_t(k);
      const int bad = !num || PyDict_SetItem(index, objective, num);
      Py_XDECREF(num);
      if (bad) goto done;
    }
  }

  /* ---- coefficient stores, variable by variable; later duplicates of a key overwrite earlier ones (:100-117) */
  const Py_ssize_t nvars = PyList_GET_SIZE(variables);
  for (Py_ssize_t j = 0; j < nvars; j++) {
    PyObject *vkey, *coefs, *hold;
    if (unpack_pair(PyList_GET_ITEM(variables, j), &vkey, &coefs, &hold)) goto done;
    const int64_t col = j + 1;
    int rc = 0;
    if (PyDict_CheckExact(coefs) && !dict_needs_js_order(coefs)) {
      Py_ssize_t pos = 0;
      PyObject *k, *v;
      Py_INCREF(coefs);
      while (!rc && PyDict_Next(coefs, &pos, &k, &v)) rc = one_coef(&st, &R, index, col, sign, k, v);
      Py_DECREF(coefs);
    } else if (PyList_CheckExact(coefs) || PyTuple_CheckExact(coefs)) {
      rc = pairs_from_iterable(&st, &R, index, col, sign, coefs);
    } else {
      PyObject *ordered = PyObject_CallFunctionObjArgs(entries, coefs, NULL);
      rc = ordered ? pairs_from_iterable(&st, &R, index, col, sign, ordered) : -1;
      Py_XDECREF(ordered);
    }
    Py_XDECREF(hold);
    if (rc) goto done;
  }

  /* ---- RHS stores (:119-127) */
  for (Py_ssize_t k = 0; k < nkeys; k++) {
    if (R.u_off[k] >= 0 && stores_push(&st, R.u_off[k], R.upper[k])) goto done;
    if (R.l_off[k] >= 0 && stores_push(&st, R.l_off[k], -R.lower[k])) goto done;
  }
  /* ---- `x <= 1` rows of the binaries, after all constraint rows (:130-134) */
  const Py_ssize_t nbin = PyList_GET_SIZE(binary_cols);
  for (Py_ssize_t b = 0; b < nbin; b++) {
    const long long col = PyLong_AsLongLong(PyList_GET_ITEM(binary_cols, b));
    if (col == -1 && PyErr_Occurred()) goto done;
    const int64_t r = (rows + b) * width;
    if (stores_push(&st, r, 1.0) || stores_push(&st, r + col, 1.0)) goto done;
  }

  {
    PyObject *cells = PyBytes_FromStringAndSize((const char *)st.cell, st.n * (Py_ssize_t)sizeof(int32_t));
    PyObject *vals = PyBytes_FromStringAndSize((const char *)st.val, st.n * (Py_ssize_t)sizeof(double));
    if (cells && vals) result = Py_BuildValue("(OOL)", cells, vals, (long long)rows);
    Py_XDECREF(cells);
    Py_XDECREF(vals);
  }

done:
  Py_DECREF(index);
  free(st.cell);
  free(st.val);
  free(R.lower);
  free(R.upper);
  free(R.u_off);
  free(R.l_off);
  free(R.is_obj);
  return result;
}

static PyMethodDef methods[] = {
    {"build", tab_build, METH_VARARGS,
     "build(variables, constraints, objective, sign, width, binary_cols, entries) -> (cells, values, rows)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_tabfast", "native core of yalps_b200.tableau", -1, methods,
                                    NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__tabfast(void) {
  s_equal = PyUnicode_InternFromString("equal");
  s_min = PyUnicode_InternFromString("min");
  s_max = PyUnicode_InternFromString("max");
  if (!s_equal || !s_min || !s_max) return NULL;
  return PyModule_Create(&module);
}
