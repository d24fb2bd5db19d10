"""Host-side logic that needs no GPU."""


def test_backend_from_environment(monkeypatch):
    """YNET_BACKEND picks the engine of a new YNet (no GPU needed to construct the module tree)."""
    import pytest as _pytest
    from motion_style_transfer_b200.models.ynet import YNet
    kw = dict(obs_len=5, pred_len=6, segmentation_model_fp=None, encoder_channels=[8, 8, 16, 16, 16],
              decoder_channels=[16, 16, 16, 8, 8], n_waypoints=1, train_net='mosa_1', position=[0, 1, 2, 3, 4],
              network='original', n_fusion=None)
    assert YNet(**kw)._backend == 'fp32'
    monkeypatch.setenv('YNET_BACKEND', 'bf16x3')
    assert YNet(**kw)._backend == 'bf16x3'
    monkeypatch.setenv('YNET_BACKEND', 'fp16')
    with _pytest.raises(ValueError):
        YNet(**kw)
