/* hicpeaks_b200._hpfast -- host-side glue of the Python mirror, nothing on the compute path.
 *
 * The reference hands `hiccups()` its band as Python lists of per-diagonal arrays (`Diags`, `cDiags`,
 * scripts/pyHICCUPS:146-157).  The C ABI wants a table of plain pointers.  Building that table in a Python loop
 * (dtype / contiguity / length check + `.ctypes.data` per array) costs ~3 us per diagonal under the GIL --
 * 1.5 ms for a 511-diagonal band, which serialises the per-chromosome host threads and used to be the largest
 * single item of an end-to-end call.  This does the same checks through the buffer protocol in ~30 us.
 *
 * Built by __graft_entry__.build():  gcc -O2 -shared -fPIC -I<python include> hp_pyhelper.c -o ../_hpfast.so
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

/* collect(seq, count, first_len, len_step, itemsize, kind, out_addr) -> int
 *   seq        list / tuple of objects exporting a buffer
 *   count      number of leading items to look at
 *   first_len  expected element count of item 0; item i must hold first_len + i * len_step elements
 *   itemsize   expected element size in bytes
 *   kind       'i' signed integer, 'f' floating point
 *   out_addr   address of a `void*[count]` table to fill (0: only validate)
 * Returns -1 when every item conforms (C-contiguous, native byte order, right type and length), otherwise the
 * index of the first item that does not.  Never raises for a non-conforming item: the caller decides whether to
 * convert it or to raise. */
static int fmt_ok(const char* f, char kind, Py_ssize_t itemsize) {
    if (!f) return 0;
    if (*f == '<' || *f == '=' || *f == '@') ++f;
    if (strlen(f) != 1) return 0;
    if (kind == 'i') {
        if (itemsize == 4) return *f == 'i' || (*f == 'l' && sizeof(long) == 4);
        if (itemsize == 8) return *f == 'q' || (*f == 'l' && sizeof(long) == 8);
        return 0;
    }
    if (kind == 'f') return (itemsize == 8 && *f == 'd') || (itemsize == 4 && *f == 'f');
    return 0;
}

static PyObject* collect(PyObject* self, PyObject* args) {
    PyObject* seq;
    Py_ssize_t count, first_len, len_step, itemsize;
    int kind;
    unsigned long long out_addr;
    if (!PyArg_ParseTuple(args, "OnnnnCK", &seq, &count, &first_len, &len_step, &itemsize, &kind, &out_addr)) return NULL;
    if (!PyList_Check(seq) && !PyTuple_Check(seq)) return PyLong_FromLong(0);
    if (PySequence_Fast_GET_SIZE(seq) < count) return PyLong_FromSsize_t(PySequence_Fast_GET_SIZE(seq));
    void** out = (void**)(uintptr_t)out_addr;
    PyObject** items = PySequence_Fast_ITEMS(seq);
    for (Py_ssize_t i = 0; i < count; ++i) {
        Py_buffer v;
        if (PyObject_GetBuffer(items[i], &v, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
            PyErr_Clear();
            return PyLong_FromSsize_t(i);
        }
        const Py_ssize_t want = first_len + i * len_step;
        const int ok = v.itemsize == itemsize && fmt_ok(v.format, (char)kind, itemsize) && v.ndim == 1 && want >= 0 &&
                       v.len == want * itemsize;
        void* p = v.buf;
        PyBuffer_Release(&v);
        if (!ok) return PyLong_FromSsize_t(i);
        if (out) out[i] = p;
    }
    return PyLong_FromLong(-1);
}

static PyMethodDef methods[] = {
    {"collect", collect, METH_VARARGS, "validate a list of per-diagonal arrays and fill a pointer table"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_hpfast", "host glue helpers of hicpeaks_b200", -1, methods};

PyMODINIT_FUNC PyInit__hpfast(void) { return PyModule_Create(&moddef); }
