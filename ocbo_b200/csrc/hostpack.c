/* Host-side staging helper (CPython extension, no CUDA): ragged list of float64 arrays ->
 * one zero-padded [B, n_pad, width] block, written straight into the pinned staging arena.
 *
 * With R trials in flight a merged device pass stages hundreds of small problems (training
 * sets, observations, candidates, normals).  The numpy spelling is one slice assignment per
 * problem and operand (~1 us each: 2 ms of a 5 ms pass at 480 problems); here it is one
 * memcpy + memset per problem (~50 ns of overhead).  Pure data movement: every floating-point
 * operation of the path still runs in libocbo_b200.so on the GPU.  ocbo_b200/engine.py falls
 * back to the numpy loop if this module is missing or an input is not a contiguous float64
 * array. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#include <string.h>

/* pack_rows(arrays, width, n_pad, out) -> list of row counts
 * arrays: sequence of C-contiguous float64 ndarrays, array b holding rows_b * width values;
 * out: writable C-contiguous float64 ndarray of len(arrays) * n_pad * width values. */
static PyObject* pack_rows(PyObject* self, PyObject* args) {
  PyObject *seq, *out_obj;
  int width, n_pad;
  if (!PyArg_ParseTuple(args, "OiiO", &seq, &width, &n_pad, &out_obj)) return NULL;
  if (width < 1 || n_pad < 1) {
    PyErr_SetString(PyExc_ValueError, "width and n_pad must be positive");
    return NULL;
  }
  if (!PyArray_Check(out_obj)) {
    PyErr_SetString(PyExc_TypeError, "out must be an ndarray");
    return NULL;
  }
  PyArrayObject* out = (PyArrayObject*)out_obj;
  PyObject* fast = PySequence_Fast(seq, "arrays must be a sequence");
  if (!fast) return NULL;
  const Py_ssize_t B = PySequence_Fast_GET_SIZE(fast);
  const npy_intp block = (npy_intp)n_pad * width;
  if (PyArray_TYPE(out) != NPY_DOUBLE || !PyArray_IS_C_CONTIGUOUS(out) || !PyArray_ISWRITEABLE(out) ||
      PyArray_SIZE(out) != (npy_intp)B * block) {
    Py_DECREF(fast);
    PyErr_SetString(PyExc_TypeError, "out must be a writable C-contiguous float64 array of B * n_pad * width values");
    return NULL;
  }
  double* dst = (double*)PyArray_DATA(out);
  PyObject* lens = PyList_New(B);
  if (!lens) {
    Py_DECREF(fast);
    return NULL;
  }
  for (Py_ssize_t b = 0; b < B; ++b) {
    PyObject* item = PySequence_Fast_GET_ITEM(fast, b);
    if (!PyArray_Check(item) || PyArray_TYPE((PyArrayObject*)item) != NPY_DOUBLE ||
        !PyArray_IS_C_CONTIGUOUS((PyArrayObject*)item)) {
      Py_DECREF(fast);
      Py_DECREF(lens);
      PyErr_SetString(PyExc_TypeError, "every array must be a C-contiguous float64 ndarray");
      return NULL;
    }
    PyArrayObject* arr = (PyArrayObject*)item;
    const npy_intp total = PyArray_SIZE(arr);
    if (total % width != 0 || total > block) {
      Py_DECREF(fast);
      Py_DECREF(lens);
      PyErr_SetString(PyExc_ValueError, "array size is not rows * width with rows <= n_pad");
      return NULL;
    }
    double* row0 = dst + (npy_intp)b * block;
    if (total) memcpy(row0, PyArray_DATA(arr), (size_t)total * sizeof(double));
    if (total < block) memset(row0 + total, 0, (size_t)(block - total) * sizeof(double));
    PyObject* n = PyLong_FromSsize_t((Py_ssize_t)(total / width));
    if (!n) {
      Py_DECREF(fast);
      Py_DECREF(lens);
      return NULL;
    }
    PyList_SET_ITEM(lens, b, n);
  }
  Py_DECREF(fast);
  return lens;
}

static PyMethodDef Methods[] = {
    {"pack_rows", pack_rows, METH_VARARGS, "ragged float64 arrays -> zero-padded block (see module doc)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef Module = {PyModuleDef_HEAD_INIT, "_hostpack",
                                    "ragged -> padded staging copies for ocbo_b200.engine", -1, Methods};

PyMODINIT_FUNC PyInit__hostpack(void) {
  import_array();
  return PyModule_Create(&Module);
}
