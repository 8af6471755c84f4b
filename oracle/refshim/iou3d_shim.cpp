// oracle/refshim/iou3d_shim.cpp -- TEST INFRASTRUCTURE (golden-vector generation only; runs in the build container).
//
// Executes the REFERENCE'S OWN compiled CPU routine
//     int boxes_iou_bev_cpu(at::Tensor boxes_a, at::Tensor boxes_b, at::Tensor ans_iou)
// from its prebuilt extension unidistill/layers/head/det3d/generate_proposals/iou3d_nms_cuda.cpython-36m-*.so on plain
// float buffers.  The extension was built against torch 1.9 / CPython 3.6 and cannot be imported, but this one routine
// touches its arguments only through three virtual calls of the TensorImpl (sizes() = vtable slot 3, dim() = slot 5,
// is_contiguous() = slot 9 or the flag byte at +0xad; read off `objdump -d`) and through at::Tensor::data_ptr<float>(),
// which it reaches via the PLT.  So it is handed look-alike objects with that vtable, the PLT symbol is supplied here,
// and the library is dlopen-ed lazily inside a process that has torch loaded (its DT_NEEDED sonames and data symbols
// resolve against the installed torch; the mismatching function symbols are never called).
// tests/golden/make_golden.py uses this to write tests/golden/iou3d_ref.npz; nothing of it ships or runs on the GPU box.
#include <dlfcn.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

struct ArrayRef { const int64_t* p; size_t n; };
struct FakeImpl;
typedef ArrayRef (*sizes_fn)(const FakeImpl*);
struct FakeImpl {
    void** vtable;
    unsigned char pad[0x100];
    int64_t sizes[2];
    float* data;
};
static ArrayRef f_sizes(const FakeImpl* t) { return ArrayRef{t->sizes, 2}; }
static int64_t f_dim(const FakeImpl*) { return 2; }
static bool f_contig(const FakeImpl*, int) { return true; }
static void f_nop(const FakeImpl*) {}
static void* g_vtable[32];

extern "C" float* _ZNK2at6Tensor8data_ptrIfEEPT_v(void* self) {        // at::Tensor::data_ptr<float>() const
    FakeImpl* impl = *reinterpret_cast<FakeImpl**>(self);
    return impl->data;
}

static void make(FakeImpl* t, float* data, int64_t rows, int64_t cols) {
    memset(t, 0, sizeof(*t));
    for (int i = 0; i < 32; ++i) g_vtable[i] = (void*)f_nop;
    g_vtable[3] = (void*)f_sizes;       // virtual IntArrayRef sizes() const
    g_vtable[5] = (void*)f_dim;         // virtual int64_t dim() const
    g_vtable[9] = (void*)f_contig;      // virtual bool is_contiguous(MemoryFormat) const
    t->vtable = g_vtable;
    reinterpret_cast<unsigned char*>(t)[0xad] = 1;   // is_contiguous_ flag of the inlined fast path
    t->sizes[0] = rows; t->sizes[1] = cols; t->data = data;
}

extern "C" int ud_refshim_boxes_iou_bev_cpu(const char* so_path, float* a, int na, float* b, int nb, float* out) {
    void* h = dlopen(so_path, RTLD_LAZY | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return -1; }
    typedef int (*fn_t)(void*, void*, void*);
    fn_t fn = (fn_t)dlsym(h, "_Z17boxes_iou_bev_cpuN2at6TensorES0_S0_");
    if (!fn) { fprintf(stderr, "dlsym: %s\n", dlerror()); return -2; }
    static FakeImpl ta, tb, tc;
    make(&ta, a, na, 7); make(&tb, b, nb, 7); make(&tc, out, na, nb);
    FakeImpl* pa = &ta; FakeImpl* pb = &tb; FakeImpl* pc = &tc;      // an at::Tensor is one pointer to its impl
    return fn(&pa, &pb, &pc);
}
