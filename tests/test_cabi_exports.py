"""CPU-side checks of the drop-in boundary: libncnn_b200.so loads without a GPU, exports every entry point that
include/ncnn_cuda.h (kernel C ABI) and include/c_api.h (reference-compatible host C API, src/c_api.h) declare, the
headers are plain C (no C++ / torch types in any signature), and the product fails loudly -- instead of falling back
to a CPU path -- when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(ROOT, "ncnn_b200", "libncnn_b200.so")


def declared(header, macro):
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
    return sorted(set(re.findall(macro + r"\s+[^;(]*?\b(\w+)\s*\(", text)))


def exported():
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    return set(l.split()[-1] for l in out.splitlines() if l.strip())


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.fail("libncnn_b200.so is not built: run python -m ncnn_b200.build")
    return C.CDLL(LIB)


def test_kernel_cabi_symbols_exported(lib):
    names = declared("ncnn_cuda.h", "NCNN_CUDA_API")
    assert len(names) >= 45, names
    have = exported()
    missing = [n for n in names if n not in have]
    assert not missing, missing
    for n in names:
        getattr(lib, n)  # resolvable through dlsym as well


def test_host_capi_symbols_exported(lib):
    names = declared("c_api.h", "NCNN_C_API")
    assert len(names) >= 120, len(names)
    have = exported()
    missing = [n for n in names if n not in have]
    assert not missing, missing


def test_exported_cxx_is_only_the_plugin_api():
    """-fvisibility=hidden: besides the extern "C" surface only the C++ plugin API in namespace ncnn (Mat, Option, Layer,
    Net, Extractor, CudaMat ... -- the classes a C++ user of the reference links against) leaves the library; the
    kernel namespace (ncnn_cuda) stays internal and is reachable through the C ABI only"""
    mangled = [s for s in exported() if s.startswith("_Z")]
    out = subprocess.run(["c++filt"], input="\n".join(mangled), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    leaked = [d for d in out if "ncnn_cuda::" in d or "tc_gemm" in d]
    assert not leaked, leaked[:10]


@pytest.mark.parametrize("header", ["ncnn_cuda.h", "c_api.h"])
def test_headers_are_plain_c(header, tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "%s"\nint main(void) { return 0; }\n' % header)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", INCLUDE, str(src)])
    text = open(os.path.join(INCLUDE, header)).read()
    for banned in ("torch", "at::", "std::", "cudaStream_t "):
        assert banned not in re.sub(r"/\*.*?\*/", "", text, flags=re.S), banned


def test_reference_capi_names_match(lib):
    """every function name of the reference's src/c_api.h that the hot path needs exists here under the same name
    (the list is the one a ctypes/cgo/JNI binding of the reference would bind for load -> input -> extract)"""
    needed = """ncnn_version ncnn_allocator_create_pool_allocator ncnn_allocator_destroy ncnn_option_create ncnn_option_destroy
    ncnn_option_set_num_threads ncnn_option_set_use_vulkan_compute ncnn_option_set_use_fp16_storage ncnn_option_set_use_bf16_storage
    ncnn_mat_create ncnn_mat_create_1d ncnn_mat_create_2d ncnn_mat_create_3d ncnn_mat_create_4d ncnn_mat_create_external_3d
    ncnn_mat_create_3d_batch ncnn_mat_destroy ncnn_mat_fill_float ncnn_mat_clone ncnn_mat_reshape_1d ncnn_mat_get_dims ncnn_mat_get_w
    ncnn_mat_get_h ncnn_mat_get_d ncnn_mat_get_c ncnn_mat_get_n ncnn_mat_get_elemsize ncnn_mat_get_elempack ncnn_mat_get_cstep
    ncnn_mat_get_nstep ncnn_mat_get_data ncnn_mat_get_channel_data ncnn_mat_get_batch_data ncnn_paramdict_create ncnn_paramdict_destroy
    ncnn_paramdict_get_type ncnn_paramdict_get_int ncnn_paramdict_get_float ncnn_paramdict_get_array ncnn_paramdict_set_int
    ncnn_paramdict_set_float ncnn_paramdict_set_array ncnn_datareader_create ncnn_datareader_create_from_memory ncnn_datareader_destroy
    ncnn_modelbin_create_from_datareader ncnn_modelbin_create_from_mat_array ncnn_modelbin_destroy ncnn_layer_create
    ncnn_layer_create_by_type ncnn_layer_destroy ncnn_layer_get_name ncnn_layer_get_type ncnn_layer_get_one_blob_only
    ncnn_layer_get_support_inplace ncnn_layer_get_bottom_count ncnn_layer_get_top_count ncnn_net_create ncnn_net_destroy
    ncnn_net_get_option ncnn_net_set_option ncnn_net_register_custom_layer_by_type ncnn_net_load_param ncnn_net_load_model
    ncnn_net_load_param_memory ncnn_net_load_model_memory ncnn_net_load_model_datareader ncnn_net_clear ncnn_net_get_input_count
    ncnn_net_get_output_count ncnn_net_get_input_name ncnn_net_get_output_name ncnn_extractor_create ncnn_extractor_destroy
    ncnn_extractor_set_option ncnn_extractor_input ncnn_extractor_extract ncnn_extractor_input_index ncnn_extractor_extract_index""".split()
    have = exported()
    missing = [n for n in needed if n not in have]
    assert not missing, missing


def test_no_cpu_fallback_without_device(lib):
    """without a CUDA device the product refuses to load a model (it must not silently compute on the CPU)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import modelzoo
    from ncnn_b200 import capi
    L = capi.library()
    L.lib.ncnn_get_cuda_device_count.restype = C.c_int
    assert L.lib.ncnn_get_cuda_device_count() == 0
    text = modelzoo.param_text("squeezenet_v1_1")
    weights = modelzoo.random_model_bytes(text, seed=1)
    opt = L.make_option(1)
    with pytest.raises(RuntimeError):
        capi.Net(L, text, weights, opt)
    L.lib.ncnn_option_destroy(opt)
    from ncnn_b200 import runner
    with pytest.raises(RuntimeError):
        runner.Session(text, weights)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under ncnn_b200/ may import, link or execute it"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ncnn_b200")):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".h", ".cu", ".cuh")):
                t = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"\boracle\b", t) and re.search(r"import oracle|from oracle|oracle/_ref|oracle\.ref", t):
                    if f != "capi.py":  # capi.py only mentions the oracle in a docstring
                        bad.append(os.path.join(dirpath, f))
                    elif re.search(r"^\s*(import oracle|from oracle)", t, re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    out = subprocess.check_output(["ldd", LIB], text=True)
    assert "ncnn_ref" not in out
