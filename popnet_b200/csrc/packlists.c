/* Host-side packer of the evaluator boundary: ragged Python lists -> CSR arrays.
 *
 * The reference's evaluator takes lists of lists (util/eval_pck.py:20-77, util/eval_mAP.py:60-157; the driver builds
 * them from JSON, main_evaluate_mp_human_3D.py:20-56): human_set[frame][human][joint][coord].  The kernels take CSR
 * (include/popnet_b200.h, PopnetPckArgs / PopnetMapArgs).  np.asarray on such nested lists costs ~100 ns per number and
 * was 85 % of the public evaluate() call; this walker does the same conversion at list-access speed.
 *
 *   pack_humans(human_set, K, D) -> (bytearray flat  [S*K*D float64], bytearray off [(N+1) int32])
 *   pack_rows(row_set, K)        -> bytearray flat  [S*K float64]            (confidence / visibility rows)
 *
 * Numbers may be Python floats, ints or anything with __float__ (NumPy scalars); containers any sequence.
 * Raises ValueError when a human is not K x D.  CPython C API only; no NumPy headers needed.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static int as_double(PyObject* o, double* out) {
  if (PyFloat_CheckExact(o)) { *out = PyFloat_AS_DOUBLE(o); return 0; }
  double v = PyFloat_AsDouble(o);
  if (v == -1.0 && PyErr_Occurred()) return -1;
  *out = v;
  return 0;
}

/* total number of second-level items (humans) and the per-frame offsets */
static Py_ssize_t count_rows(PyObject* outer_fast, int32_t* off) {
  const Py_ssize_t N = PySequence_Fast_GET_SIZE(outer_fast);
  Py_ssize_t total = 0;
  if (off) off[0] = 0;
  for (Py_ssize_t f = 0; f < N; ++f) {
    PyObject* fr = PySequence_Fast_GET_ITEM(outer_fast, f);
    const Py_ssize_t n = PySequence_Size(fr);
    if (n < 0) return -1;
    total += n;
    if (total > INT32_MAX) { PyErr_SetString(PyExc_OverflowError, "too many humans for int32 offsets"); return -1; }
    if (off) off[f + 1] = (int32_t)total;
  }
  return total;
}

/* one row of K numbers (D == 0) or K joints of D numbers */
static int fill_row(PyObject* row, Py_ssize_t K, Py_ssize_t D, double* dst) {
  PyObject* rf = PySequence_Fast(row, "a human / row must be a sequence");
  if (!rf) return -1;
  if (PySequence_Fast_GET_SIZE(rf) != K) {
    PyErr_Format(PyExc_ValueError, "every human must have %zd joints, got %zd", K, PySequence_Fast_GET_SIZE(rf));
    Py_DECREF(rf);
    return -1;
  }
  for (Py_ssize_t k = 0; k < K; ++k) {
    PyObject* j = PySequence_Fast_GET_ITEM(rf, k);
    if (D == 0) {
      if (as_double(j, dst + k) < 0) { Py_DECREF(rf); return -1; }
      continue;
    }
    PyObject* jf = PySequence_Fast(j, "a joint must be a sequence of coordinates");
    if (!jf) { Py_DECREF(rf); return -1; }
    if (PySequence_Fast_GET_SIZE(jf) != D) {
      PyErr_Format(PyExc_ValueError, "every joint must have %zd coordinates, got %zd", D, PySequence_Fast_GET_SIZE(jf));
      Py_DECREF(jf); Py_DECREF(rf);
      return -1;
    }
    for (Py_ssize_t d = 0; d < D; ++d)
      if (as_double(PySequence_Fast_GET_ITEM(jf, d), dst + k * D + d) < 0) { Py_DECREF(jf); Py_DECREF(rf); return -1; }
    Py_DECREF(jf);
  }
  Py_DECREF(rf);
  return 0;
}

static PyObject* pack_impl(PyObject* set, Py_ssize_t K, Py_ssize_t D, int want_off) {
  PyObject* outer = PySequence_Fast(set, "expected a sequence of frames");
  if (!outer) return NULL;
  const Py_ssize_t N = PySequence_Fast_GET_SIZE(outer);
  PyObject* off_ba = NULL;
  int32_t* off = NULL;
  if (want_off) {
    off_ba = PyByteArray_FromStringAndSize(NULL, (N + 1) * (Py_ssize_t)sizeof(int32_t));
    if (!off_ba) { Py_DECREF(outer); return NULL; }
    off = (int32_t*)PyByteArray_AS_STRING(off_ba);
  }
  const Py_ssize_t S = count_rows(outer, off);
  if (S < 0) { Py_XDECREF(off_ba); Py_DECREF(outer); return NULL; }
  const Py_ssize_t per = K * (D == 0 ? 1 : D);
  PyObject* flat_ba = PyByteArray_FromStringAndSize(NULL, S * per * (Py_ssize_t)sizeof(double));
  if (!flat_ba) { Py_XDECREF(off_ba); Py_DECREF(outer); return NULL; }
  double* dst = (double*)PyByteArray_AS_STRING(flat_ba);
  for (Py_ssize_t f = 0; f < N; ++f) {
    PyObject* fr = PySequence_Fast(PySequence_Fast_GET_ITEM(outer, f), "a frame must be a sequence of humans");
    if (!fr) goto fail;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fr);
    for (Py_ssize_t h = 0; h < n; ++h) {
      if (fill_row(PySequence_Fast_GET_ITEM(fr, h), K, D, dst) < 0) { Py_DECREF(fr); goto fail; }
      dst += per;
    }
    Py_DECREF(fr);
  }
  Py_DECREF(outer);
  if (!want_off) return flat_ba;
  {
    PyObject* res = PyTuple_Pack(2, flat_ba, off_ba);
    Py_DECREF(flat_ba); Py_DECREF(off_ba);
    return res;
  }
fail:
  Py_DECREF(flat_ba); Py_XDECREF(off_ba); Py_DECREF(outer);
  return NULL;
}

static PyObject* py_pack_humans(PyObject* self, PyObject* args) {
  PyObject* set; Py_ssize_t K, D;
  if (!PyArg_ParseTuple(args, "Onn", &set, &K, &D)) return NULL;
  if (K < 1 || D < 1) { PyErr_SetString(PyExc_ValueError, "K and D must be positive"); return NULL; }
  return pack_impl(set, K, D, 1);
}

static PyObject* py_pack_rows(PyObject* self, PyObject* args) {
  PyObject* set; Py_ssize_t K;
  if (!PyArg_ParseTuple(args, "On", &set, &K)) return NULL;
  if (K < 1) { PyErr_SetString(PyExc_ValueError, "K must be positive"); return NULL; }
  return pack_impl(set, K, 0, 0);
}

static PyMethodDef methods[] = {
    {"pack_humans", py_pack_humans, METH_VARARGS, "pack_humans(human_set, K, D) -> (flat float64 bytearray, off int32 bytearray)"},
    {"pack_rows", py_pack_rows, METH_VARARGS, "pack_rows(row_set, K) -> flat float64 bytearray"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_packlists", "ragged lists -> CSR packer of the evaluator boundary", -1, methods};

PyMODINIT_FUNC PyInit__packlists(void) { return PyModule_Create(&moddef); }
