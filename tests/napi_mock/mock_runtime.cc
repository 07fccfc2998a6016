// Minimal N-API runtime for tests: implements the functions declared in the mock node_api.h over a tagged value
// type, loads the addon's exports through its NAPI_MODULE registration and lets a ctypes test call them the way
// Node would (tests/test_napi_addon.py).  Semantics follow the Node documentation for each function used:
// napi_get_cb_info reports the actual argument count and pads with undefined, typed-array / buffer accessors
// return napi_invalid_arg for other kinds, napi_throw_* leaves a pending exception on the env.
#include <node_api.h>

#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

enum Kind { K_UNDEFINED, K_NUMBER, K_EXTERNAL, K_BUFFER, K_TYPEDARRAY, K_FUNCTION, K_OBJECT };
struct napi_value__ {
    Kind kind = K_UNDEFINED;
    double num = 0;
    void* ptr = nullptr;
    size_t len = 0;
    napi_typedarray_type tt = napi_uint8_array;
    napi_callback cb = nullptr;
    napi_finalize fin = nullptr; void* fin_hint = nullptr;      // externals: run when the env is destroyed (garbage collection)
    std::string name;
    std::vector<std::pair<std::string, napi_value>> props;
};
struct napi_env__ {
    std::vector<std::unique_ptr<napi_value__>> heap;
    bool pending = false;
    std::string err_kind, err_code, err_msg, last;
    napi_value exports = nullptr;
    napi_value make(Kind k) { heap.emplace_back(new napi_value__); heap.back()->kind = k; return heap.back().get(); }
};
struct napi_callback_info__ { std::vector<napi_value> args; };

extern "C" {
napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
    if (!env || !info) return napi_invalid_arg;
    if (argc && argv) {
        const size_t cap = *argc;
        for (size_t i = 0; i < cap; ++i) argv[i] = i < info->args.size() ? info->args[i] : env->make(K_UNDEFINED);
    }
    if (argc) *argc = info->args.size();
    if (this_arg) *this_arg = env->exports;
    if (data) *data = nullptr;
    return napi_ok;
}
napi_status napi_get_buffer_info(napi_env, napi_value v, void** data, size_t* length) {
    if (!v || v->kind != K_BUFFER) return napi_invalid_arg;
    if (data) *data = v->ptr;
    if (length) *length = v->len;
    return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env env, napi_value v, napi_typedarray_type* type, size_t* length, void** data, napi_value* ab, size_t* off) {
    if (!v || v->kind != K_TYPEDARRAY) return napi_invalid_arg;
    if (type) *type = v->tt;
    if (length) *length = v->len;          // element count, as in Node
    if (data) *data = v->ptr;
    if (ab) *ab = env->make(K_OBJECT);
    if (off) *off = 0;
    return napi_ok;
}
napi_status napi_get_value_int32(napi_env, napi_value v, int32_t* r) { if (!v || v->kind != K_NUMBER) return napi_number_expected; *r = (int32_t)(int64_t)v->num; return napi_ok; }
napi_status napi_get_value_int64(napi_env, napi_value v, int64_t* r) { if (!v || v->kind != K_NUMBER) return napi_number_expected; *r = (int64_t)v->num; return napi_ok; }
napi_status napi_get_value_double(napi_env, napi_value v, double* r) { if (!v || v->kind != K_NUMBER) return napi_number_expected; *r = v->num; return napi_ok; }
napi_status napi_get_value_external(napi_env, napi_value v, void** r) { if (!v || v->kind != K_EXTERNAL) return napi_invalid_arg; *r = v->ptr; return napi_ok; }
napi_status napi_create_external(napi_env env, void* data, napi_finalize fin, void* hint, napi_value* r) {
    *r = env->make(K_EXTERNAL); (*r)->ptr = data; (*r)->fin = fin; (*r)->fin_hint = hint; return napi_ok;
}
napi_status napi_create_int32(napi_env env, int32_t x, napi_value* r) { *r = env->make(K_NUMBER); (*r)->num = x; return napi_ok; }
napi_status napi_create_function(napi_env env, const char* name, size_t, napi_callback cb, void*, napi_value* r) {
    *r = env->make(K_FUNCTION); (*r)->cb = cb; (*r)->name = name ? name : ""; return napi_ok;
}
napi_status napi_set_named_property(napi_env, napi_value obj, const char* name, napi_value v) {
    if (!obj || obj->kind != K_OBJECT) return napi_object_expected;
    obj->props.emplace_back(name, v); return napi_ok;
}
static napi_status throw_kind(napi_env env, const char* kind, const char* code, const char* msg) {
    env->pending = true; env->err_kind = kind; env->err_code = code ? code : ""; env->err_msg = msg ? msg : ""; return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char* code, const char* msg) { return throw_kind(env, "Error", code, msg); }
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg) { return throw_kind(env, "TypeError", code, msg); }
napi_status napi_throw_range_error(napi_env env, const char* code, const char* msg) { return throw_kind(env, "RangeError", code, msg); }

// ---- driver API for the ctypes test --------------------------------------------------------------------
napi_env mock_env_create(void) {
    napi_env env = new napi_env__;
    env->exports = env->make(K_OBJECT);
    napi_value r = napi_mock_module_init(env, env->exports);      // what `require('./jsrt_addon.node')` triggers
    if (r && r->kind == K_OBJECT) env->exports = r;
    return env;
}
// tearing the environment down collects every value: externals run their finalizers, like Node's garbage collector would
void mock_env_destroy(napi_env env) {
    for (auto& v : env->heap) if (v->kind == K_EXTERNAL && v->fin) { v->fin(env, v->ptr, v->fin_hint); v->fin = nullptr; }
    delete env;
}
int mock_run_finalizer(napi_env env, napi_value v) {          // collect one external now
    if (!v || v->kind != K_EXTERNAL || !v->fin) return 0;
    v->fin(env, v->ptr, v->fin_hint); v->fin = nullptr; v->ptr = nullptr; return 1;
}
int mock_export_count(napi_env env) { return (int)env->exports->props.size(); }
const char* mock_export_name(napi_env env, int i) { return env->exports->props[i].first.c_str(); }
napi_value mock_number(napi_env env, double x) { napi_value v = env->make(K_NUMBER); v->num = x; return v; }
napi_value mock_buffer(napi_env env, void* p, size_t len) { napi_value v = env->make(K_BUFFER); v->ptr = p; v->len = len; return v; }
napi_value mock_typedarray(napi_env env, int type, void* p, size_t n) { napi_value v = env->make(K_TYPEDARRAY); v->tt = (napi_typedarray_type)type; v->ptr = p; v->len = n; return v; }
// calls exports[name](...argv); returns the result (NULL = undefined).  A thrown exception is left pending.
napi_value mock_call(napi_env env, const char* name, int argc, napi_value* argv) {
    for (auto& p : env->exports->props)
        if (p.first == name && p.second->kind == K_FUNCTION) {
            napi_callback_info__ info; info.args.assign(argv, argv + argc);
            return p.second->cb(env, &info);
        }
    throw_kind(env, "TypeError", "", "not a function");
    return nullptr;
}
// "Kind: message" of the pending exception (and clears it), or NULL
const char* mock_take_exception(napi_env env) {
    if (!env->pending) return nullptr;
    env->pending = false;
    env->last = env->err_kind + ": " + env->err_msg;
    return env->last.c_str();
}
int mock_is_number(napi_value v) { return v && v->kind == K_NUMBER; }
int mock_is_external(napi_value v) { return v && v->kind == K_EXTERNAL; }
double mock_number_value(napi_value v) { return v ? v->num : 0; }
}
