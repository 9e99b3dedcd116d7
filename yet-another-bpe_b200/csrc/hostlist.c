/* hostlist.c -- the host tail of train_bpe: Python objects of the result, built in C.
 *
 * The reference API returns a dict of vocab_size entries and a list of (bytes, bytes) tuples (trainer.py:94-134, 296-300;
 * tests/adapters.py:66-99).  With 32 000 merges that is ~65 000 small objects: ~25 ms of CPython bytecode (comprehensions), 10 % of
 * the whole B200 training step.  This module builds the same objects with the C API in one call.  Host glue only: no compute,
 * nothing here touches the device; yabpe falls back to the Python construction when the module is not built.
 *
 *   materialise(pool: bytes-like, offs: int64 buffer [ntok + 1], merges: int32 buffer [2 * n_merges])
 *       -> (tokens: list[bytes], vocab: dict[bytes, int], merges: list[tuple[bytes, bytes]])
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static PyObject* materialise(PyObject* self, PyObject* args) {
    Py_buffer pool, offs, mg;
    if (!PyArg_ParseTuple(args, "y*y*y*", &pool, &offs, &mg)) return NULL;
    PyObject *tokens = NULL, *vocab = NULL, *merges = NULL, *ret = NULL;
    const Py_ssize_t ntok = offs.len / (Py_ssize_t)sizeof(int64_t) - 1, nm = mg.len / (Py_ssize_t)(2 * sizeof(int32_t));
    const int64_t* o = (const int64_t*)offs.buf;
    const int32_t* m = (const int32_t*)mg.buf;
    const char* p = (const char*)pool.buf;
    if (ntok < 0 || offs.len % sizeof(int64_t) || mg.len % (2 * sizeof(int32_t))) { PyErr_SetString(PyExc_ValueError, "bad buffer sizes"); goto done; }
    tokens = PyList_New(ntok);
    vocab = PyDict_New();
    merges = PyList_New(nm);
    if (!tokens || !vocab || !merges) goto done;
    for (Py_ssize_t i = 0; i < ntok; i++) {
        if (o[i] < 0 || o[i + 1] < o[i] || o[i + 1] > (int64_t)pool.len) { PyErr_SetString(PyExc_ValueError, "token offsets out of range"); goto done; }
        PyObject* b = PyBytes_FromStringAndSize(p + o[i], (Py_ssize_t)(o[i + 1] - o[i]));
        if (!b) goto done;
        PyList_SET_ITEM(tokens, i, b);                       /* steals the reference */
        PyObject* id = PyLong_FromSsize_t(i);
        if (!id) goto done;
        const int rc = PyDict_SetItem(vocab, b, id);         /* equal bytes: the later id wins, as in {b: i for i, b in enumerate(tokens)} */
        Py_DECREF(id);
        if (rc < 0) goto done;
    }
    for (Py_ssize_t k = 0; k < nm; k++) {
        const int32_t a = m[2 * k], b = m[2 * k + 1];
        if (a < 0 || b < 0 || a >= ntok || b >= ntok) { PyErr_SetString(PyExc_ValueError, "merge refers to an unknown token"); goto done; }
        PyObject* t = PyTuple_Pack(2, PyList_GET_ITEM(tokens, a), PyList_GET_ITEM(tokens, b));
        if (!t) goto done;
        PyList_SET_ITEM(merges, k, t);
    }
    ret = PyTuple_Pack(3, tokens, vocab, merges);
done:
    Py_XDECREF(tokens); Py_XDECREF(vocab); Py_XDECREF(merges);
    PyBuffer_Release(&pool); PyBuffer_Release(&offs); PyBuffer_Release(&mg);
    return ret;
}

static PyMethodDef methods[] = {
    {"materialise", materialise, METH_VARARGS, "tokens, vocab and merges of a training result as Python objects"},
    {NULL, NULL, 0, NULL},
};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_hostlist", "host tail of yabpe.train_bpe", -1, methods};
PyMODINIT_FUNC PyInit__hostlist(void) { return PyModule_Create(&moddef); }
