"""Host side of the training paths (video2music_b200.autograd, the module-level dispatch in moe.py / mamba.py / rpr.py /
grouped_query_attention.py / custom_transformer.py / video_regression.py) without a GPU: the GPU parity tests of
tests/test_gpu_variant_train.py are re-run with `video2music_b200.ops` replaced by the CPU mirrors of tests/kernel_mirror.py.
What this proves on a CPU box: the autograd Functions route every gradient to the right parameter with the right strides and
scalings, and the chunk / carry algebra of the scan and conv backward kernels is right -- all against the gradients of the
unmodified reference stored in tests/golden/.  The kernels themselves are only proven by the `-m gpu` run of the same tests."""
import pytest

import kernel_mirror
import test_gpu_variant_train as T


@pytest.fixture
def mirrored(monkeypatch):
    kernel_mirror.install(monkeypatch)
    monkeypatch.setattr(T, "DEV", "cpu")


@pytest.mark.parametrize("shared", [False, True])
def test_moe_layer_gradients_host(mirrored, shared):
    T.test_moe_train_golden_gpu(shared)


@pytest.mark.parametrize("name", ["post_ln_moe", "post_ln_sharedmoe_b2", "pre_rms_moe"])
def test_gqa_moe_stack_gradients_host(mirrored, name):
    T.test_variant_train_golden_gpu(name)


def test_gqa_function_gradients_host(mirrored):
    T.test_gqa_function_backward_vs_oracle(2, 33, 33, 8, 2, True)
    T.test_gqa_function_backward_vs_oracle(3, 20, 45, 8, 4, False)


@pytest.mark.parametrize("name", ["block_v0", "block_v1", "stack", "bimamba_layer", "bimamba_v1_ffn", "bimamba_v1_moe"])
def test_mamba_family_gradients_host(mirrored, name):
    T.test_mamba_train_golden_gpu(name)


@pytest.mark.parametrize("B,L,ED,plus", [(2, 33, 100, True), (1, 200, 40, True), (2, 70, 64, False)])
def test_scan_and_conv_backward_mirror_vs_autograd(mirrored, B, L, ED, plus):
    """The step-by-step mirrors of the chunked scan / conv backward kernels against float64 autograd of the recurrence."""
    T.test_selective_scan_and_conv_backward_kernels_vs_autograd(B, L, ED, plus)


@pytest.mark.parametrize("reg", ["mamba+", "bimamba+", "sharedmoe_bimamba+"])
def test_video_regression_gradients_host(mirrored, reg):
    T.test_video_regression_train_golden_gpu(reg)


def test_rpr_module_and_decoder_gradients_host(mirrored):
    T.test_rpr_module_backward_golden_gpu()
    T.test_rpr_decoder_train_golden_gpu("layer")
    T.test_rpr_decoder_train_golden_gpu("decoder")


@pytest.mark.parametrize("ver", ["2.2", "1.1", "3.1"])
def test_zoo_model_training_host(mirrored, ver):
    """Host side of the model-zoo training paths (custom / differential attention autograd, RoPE backward, embeddings, position tables)
    against the unmodified reference's gradients (tests/golden/zoo_train.pt), ops replaced by the CPU mirrors."""
    T.test_zoo_model_train_step_vs_reference_golden(ver)


def test_mirrors_are_not_installed_outside_the_fixture():
    """The product path has no CPU fallback: without the fixture the same call raises."""
    import torch
    from video2music_b200 import GLUExpert
    with pytest.raises(Exception):
        GLUExpert(8, 16, 0.0)(torch.zeros(2, 8, requires_grad=True))
