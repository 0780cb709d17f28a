"""Pins the oracle's DINOv2 restatement (un-vendored transformers==4.50.0 FlaxDinov2Module) against the
torch `transformers.Dinov2Model` available locally: same published architecture, weights copied across.
The position table is pre-interpolated by the oracle so the (unpinned) bicubic step is excluded."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
transformers = pytest.importorskip("transformers")


def _load(model, dino):
    sd = model.state_dict()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    emb = dino["embeddings"]
    sd["embeddings.cls_token"].copy_(t(emb["cls_token"]))
    k = emb["patch_embeddings"]["projection"]["kernel"]                  # (kh,kw,cin,cout) -> (cout,cin,kh,kw)
    sd["embeddings.patch_embeddings.projection.weight"].copy_(t(k.transpose(3, 2, 0, 1)))
    sd["embeddings.patch_embeddings.projection.bias"].copy_(t(emb["patch_embeddings"]["projection"]["bias"]))
    for l in range(12):
        L = dino["encoder"]["layer"][str(l)]
        p = f"encoder.layer.{l}."
        a = L["attention"]["attention"]
        for nm in ("query", "key", "value"):
            sd[p + f"attention.attention.{nm}.weight"].copy_(t(a[nm]["kernel"].T))
            sd[p + f"attention.attention.{nm}.bias"].copy_(t(a[nm]["bias"]))
        sd[p + "attention.output.dense.weight"].copy_(t(L["attention"]["output"]["dense"]["kernel"].T))
        sd[p + "attention.output.dense.bias"].copy_(t(L["attention"]["output"]["dense"]["bias"]))
        sd[p + "layer_scale1.lambda1"].copy_(t(L["layer_scale1"]["lambda1"]))
        sd[p + "layer_scale2.lambda1"].copy_(t(L["layer_scale2"]["lambda1"]))
        for nm in ("norm1", "norm2"):
            sd[p + nm + ".weight"].copy_(t(L[nm]["scale"]))
            sd[p + nm + ".bias"].copy_(t(L[nm]["bias"]))
        for nm in ("fc1", "fc2"):
            sd[p + f"mlp.{nm}.weight"].copy_(t(L["mlp"][nm]["kernel"].T))
            sd[p + f"mlp.{nm}.bias"].copy_(t(L["mlp"][nm]["bias"]))
    sd["layernorm.weight"].copy_(t(dino["layernorm"]["scale"]))
    sd["layernorm.bias"].copy_(t(dino["layernorm"]["bias"]))
    model.load_state_dict(sd)


def test_oracle_dinov2_matches_torch_dinov2(params_p1):
    from hvla import params as P, synthetic as S
    from oracle import hypervla_oracle as O
    dino = P.dino_tree_from_params(params_p1)
    cfg = transformers.Dinov2Config(image_size=518, patch_size=14)
    assert (cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.hidden_act) == (768, 12, 12, "gelu")
    model = transformers.Dinov2Model(cfg).eval()
    _load(model, dino)
    images = S.make_inputs(2, 2, 2)["images"][:, 0]
    pos = O.interpolate_pos_table(dino["embeddings"]["position_embeddings"])
    mine = O.dinov2_forward(dino, images, np.float32, pos)
    with torch.no_grad():
        # patch embedding (conv) by torch on the same normalised NCHW pixels
        x = images.astype(np.float32) / 255.0
        x = (x - np.asarray(O.IMAGE_MEAN, np.float32)) / np.asarray(O.IMAGE_STD, np.float32)
        px = model.embeddings.patch_embeddings(torch.from_numpy(x.transpose(0, 3, 1, 2).copy()))
        mine_emb = O.dinov2_embed(dino, images, pos, np.float32)
        ref_emb = torch.cat([model.embeddings.cls_token.expand(2, -1, -1), px], 1) + torch.from_numpy(pos)[None]
        assert np.abs(mine_emb - ref_emb.numpy()).max() < 2e-5
        out = model.encoder(torch.from_numpy(mine_emb))
        hs = out.last_hidden_state if hasattr(out, "last_hidden_state") else out[0]
        ref = model.layernorm(hs).numpy()
    err = np.abs(mine - ref).max() / np.abs(ref).max()
    assert err < 2e-5, err
