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


def test_split_weight_gather_index_equals_block_construction():
    """ops._split_weight_select: the cached gather index that lays [W_hi | W_lo | 0] out for ynet_tc_conv3x3_split equals
    the block-by-block construction (per source [hi | hi] over its stored channels, then lo; every part padded to its cp),
    bit for bit, on random layouts -- fine-tuning re-packs the adapted layers' weights after every optimiser step."""
    import random
    import pytest as _pytest
    import torch
    from motion_style_transfer_b200 import ops

    def blocks(w, src_layouts):
        C_out, _, kh, kw = w.shape
        w_hi = w.to(torch.bfloat16).to(torch.float32)
        w_lo = w - w_hi
        out, chans, c0 = [], [], 0
        for layout in src_layouts:
            tot = sum(cp for _, cp in layout)
            a, b, o = torch.zeros(C_out, 2 * tot, kh, kw), torch.zeros(C_out, tot, kh, kw), 0
            for C, cp in layout:
                a[:, o:o + C] = w_hi[:, c0:c0 + C]
                a[:, tot + o:tot + o + C] = w_hi[:, c0:c0 + C]
                b[:, o:o + C] = w_lo[:, c0:c0 + C]
                o += cp
                c0 += C
            out += [a, b]
            chans += [2 * tot, tot]
        return torch.cat(out, 1), chans

    rnd = random.Random(0)
    torch.manual_seed(0)
    for _ in range(60):
        layouts = [[(C, -(-C // 16) * 16) for C in [rnd.randint(1, 40) for _ in range(rnd.randint(1, 3))]]
                   for _ in range(rnd.randint(1, 2))]
        c_in = sum(C for layout in layouts for C, _ in layout)
        w = torch.randn(rnd.randint(1, 9), c_in, *rnd.choice([(1, 1), (3, 3)]))
        ref, chans = blocks(w, layouts)
        sel, chans2 = ops._split_weight_select(layouts, c_in, 'cpu')
        w_hi = w.to(torch.bfloat16).to(torch.float32)
        got = torch.cat([w_hi, w - w_hi, w.new_zeros(w.shape[0], 1, *w.shape[2:])], 1).index_select(1, sel)
        assert chans == chans2 and torch.equal(ref, got) and got.is_contiguous()
        assert ops._split_weight_select(layouts, c_in, 'cpu')[0] is sel          # cached
    with _pytest.raises(ValueError, match='sources hold 3 channels'):
        ops._split_weight_select([[(3, 16)]], 4, 'cpu')
