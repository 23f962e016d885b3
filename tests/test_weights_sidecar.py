"""Side-car persistence of the converted weights, host logic only: the bundle's state round-trips through
torch.save / torch.load(weights_only=True), the source fingerprint follows the bytes, format / geometry mismatches are refused."""
import pytest
import torch

from nunchaku_b200.weights import B200Weights


def _bundle():
    g = torch.Generator().manual_seed(0)
    N, K, R = 128, 256, 32
    return B200Weights(N=N, K=K, rank=R, fp4=True, dtype=torch.bfloat16,
                       qweight=torch.randint(0, 255, (N, K // 2), generator=g, dtype=torch.uint8),
                       wscales=torch.randint(0, 120, (N * K // 16,), generator=g, dtype=torch.uint8),
                       bias=torch.randn(N, generator=g), cscale=torch.rand(N, generator=g) + 0.5,
                       lora_up=torch.randn(N * R, generator=g).to(torch.bfloat16), lora_down=torch.randn(2 * K * R, generator=g).to(torch.bfloat16),
                       lora_down_next=torch.randn(R, K, generator=g).to(torch.bfloat16), smooth=None)


def test_state_round_trip(tmp_path):
    w = _bundle()
    path = tmp_path / "layer.b200"
    torch.save(w.state(), path)
    w2 = B200Weights.from_state(torch.load(path, map_location="cpu", weights_only=True), "cpu")
    assert (w2.N, w2.K, w2.rank, w2.fp4, w2.dtype) == (w.N, w.K, w.rank, w.fp4, w.dtype) and w2.smooth is None
    for name in ("qweight", "wscales", "bias", "cscale", "lora_up", "lora_down", "lora_down_next"):
        a, b = getattr(w, name), getattr(w2, name)
        assert a.dtype == b.dtype and torch.equal(a.view(torch.uint8), b.view(torch.uint8)), name
    assert w2.nbytes() == w.nbytes()


def test_refuses_foreign_files():
    w = _bundle()
    st = w.state()
    st["format"] = 99
    with pytest.raises(ValueError, match="format"):
        B200Weights.from_state(st, "cpu")
    st = w.state()
    st["K"] = 512
    with pytest.raises(ValueError, match="corrupt"):
        B200Weights.from_state(st, "cpu")


def test_fingerprint_follows_the_bytes():
    a = torch.arange(64, dtype=torch.int8)
    b = torch.ones(8, dtype=torch.bfloat16)
    f0 = B200Weights.fingerprint(qweight=a, bias=b, alpha=1.0, wcscales=None)
    assert f0 == B200Weights.fingerprint(bias=b.clone(), qweight=a.clone(), alpha=1.0, wcscales=None)   # order / identity independent
    a2 = a.clone()
    a2[17] += 1
    assert f0 != B200Weights.fingerprint(qweight=a2, bias=b, alpha=1.0, wcscales=None)
    assert f0 != B200Weights.fingerprint(qweight=a, bias=b, alpha=0.5, wcscales=None)
    assert f0 != B200Weights.fingerprint(qweight=a.view(8, 8), bias=b, alpha=1.0, wcscales=None)


def test_converted_weights_follow_a_device_move():
    """module.to(device) after the conversion moves the B200Weights bundle with the parameters (CPU stand-in: cpu -> meta); a dtype cast leaves it alone."""
    import torch

    from nunchaku_b200.models.linear import SVDQW4A4Linear
    from nunchaku_b200.weights import B200Weights

    m = SVDQW4A4Linear(128, 128, rank=16, precision="nvfp4", torch_dtype=torch.bfloat16, device="cpu")
    z = lambda *shape, dt=torch.uint8: torch.zeros(*shape, dtype=dt)   # noqa: E731
    m._b200 = B200Weights(N=128, K=128, rank=16, fp4=True, dtype=torch.bfloat16, qweight=z(128, 64), wscales=z(1024), bias=z(128, dt=torch.float32),
                          cscale=None, lora_up=z(128, 32, dt=torch.bfloat16), lora_down=z(128, 16, dt=torch.bfloat16), lora_down_next=None,
                          smooth=z(128, dt=torch.bfloat16))
    m._b200_alpha = m.wtscale
    m.to(torch.float16)                                    # cast: parameters change dtype, the bundle (fp32 bias included) is untouched
    assert m._b200.bias.dtype == torch.float32 and m._b200.qweight.device.type == "cpu"
    m.to("meta")
    assert m.qweight.device.type == "meta"
    assert all(getattr(m._b200, n).device.type == "meta" for n in ("qweight", "wscales", "bias", "lora_up", "lora_down", "smooth"))
    assert m._b200.cscale is None and m._b200.lora_down_next is None
