"""Test helper: write FunASR-style checkpoint folders (model.pt with the module-tree keys the reference exporters walk, Kaldi
am.mvn, config.yaml) from the seeded synthetic checkpoints."""
import torch


def _mvn(path, means, scales):
    row = lambda v: " ".join(f"{float(x):.9g}" for x in v)
    n = len(means)
    path.write_text(f"<Nnet>\n<Splice> {n} {n}\n[ 0 ]\n<AddShift> {n} {n}\n<LearnRateCoef> 0 [ {row(means)} ]\n"
                    f"<Rescale> {n} {n}\n<LearnRateCoef> 0 [ {row(scales)} ]\n</Nnet>\n")


def _enc_layer(sd, key, raw, p, D, k):
    sd[key + "norm1.weight"], sd[key + "norm1.bias"] = raw[p + "norm1.g"], raw[p + "norm1.b"]
    sd[key + "norm2.weight"], sd[key + "norm2.bias"] = raw[p + "norm2.g"], raw[p + "norm2.b"]
    sd[key + "self_attn.linear_q_k_v.weight"], sd[key + "self_attn.linear_q_k_v.bias"] = raw[p + "qkv.w"], raw[p + "qkv.b"]
    sd[key + "self_attn.linear_out.weight"], sd[key + "self_attn.linear_out.bias"] = raw[p + "out.w"], raw[p + "out.b"]
    sd[key + "self_attn.fsmn_block.weight"] = raw[p + "fsmn.w"].reshape(D, 1, k)
    sd[key + "feed_forward.w_1.weight"], sd[key + "feed_forward.w_1.bias"] = raw[p + "w1.w"], raw[p + "w1.b"]
    sd[key + "feed_forward.w_2.weight"], sd[key + "feed_forward.w_2.bias"] = raw[p + "w2.w"], raw[p + "w2.b"]


def write_sensevoice_folder(folder, d, raw):
    sd = {"embed.weight": raw["embed"], "ctc.ctc_lo.weight": raw["ctc.w"], "ctc.ctc_lo.bias": raw["ctc.b"]}
    names = [f"encoder.encoders0.{i}." for i in range(d.n_blocks0)] + [f"encoder.encoders.{i}." for i in range(d.n_blocks)] + \
            [f"encoder.tp_encoders.{i}." for i in range(d.n_tp_blocks)]
    for i, key in enumerate(names):
        _enc_layer(sd, key, raw, f"blk{i}.", d.d_model, d.fsmn_kernel)
    for n, key in (("after_norm", "encoder.after_norm"), ("tp_norm", "encoder.tp_norm")):
        sd[key + ".weight"], sd[key + ".bias"] = raw[n + ".g"], raw[n + ".b"]
    torch.save(sd, folder / "model.pt")
    _mvn(folder / "am.mvn", raw["cmvn_means"], raw["cmvn_vars"])
    (folder / "config.yaml").write_text(f"encoder_conf:\n  attention_heads: {d.n_heads}\n  output_size: {d.d_model}\n")


def write_paraformer_folder(folder, d, raw):
    sd = {}
    names = [f"encoder.encoders0.{i}." for i in range(d.n_blocks0)] + [f"encoder.encoders.{i}." for i in range(d.n_blocks)]
    for i, key in enumerate(names):
        _enc_layer(sd, key, raw, f"enc{i}.", d.d_model, d.fsmn_kernel)
    sd["encoder.after_norm.weight"], sd["encoder.after_norm.bias"] = raw["enc_after_norm.g"], raw["enc_after_norm.b"]
    sd["predictor.cif_conv1d.weight"], sd["predictor.cif_conv1d.bias"] = raw["cif.conv.w"], raw["cif.conv.b"]
    sd["predictor.cif_output.weight"], sd["predictor.cif_output.bias"] = raw["cif.out.w"], raw["cif.out.b"]
    dec = [f"decoder.decoders.{i}." for i in range(d.dec_att_blocks)] + [f"decoder.decoders3.{i}." for i in range(d.dec_ffn_blocks)]
    for i, key in enumerate(dec):
        p = f"dec{i}."
        sd[key + "norm1.weight"], sd[key + "norm1.bias"] = raw[p + "norm1.g"], raw[p + "norm1.b"]
        sd[key + "feed_forward.norm.weight"], sd[key + "feed_forward.norm.bias"] = raw[p + "ffn_norm.g"], raw[p + "ffn_norm.b"]
        sd[key + "feed_forward.w_1.weight"], sd[key + "feed_forward.w_1.bias"] = raw[p + "w1.w"], raw[p + "w1.b"]
        sd[key + "feed_forward.w_2.weight"] = raw[p + "w2.w"]
        if i < d.dec_att_blocks:
            sd[key + "norm2.weight"], sd[key + "norm2.bias"] = raw[p + "norm2.g"], raw[p + "norm2.b"]
            sd[key + "norm3.weight"], sd[key + "norm3.bias"] = raw[p + "norm3.g"], raw[p + "norm3.b"]
            sd[key + "self_attn.fsmn_block.weight"] = raw[p + "fsmn.w"].reshape(d.d_model, 1, d.fsmn_kernel)
            sd[key + "src_attn.linear_q.weight"], sd[key + "src_attn.linear_q.bias"] = raw[p + "q.w"], raw[p + "q.b"]
            sd[key + "src_attn.linear_k_v.weight"], sd[key + "src_attn.linear_k_v.bias"] = raw[p + "kv.w"], raw[p + "kv.b"]
            sd[key + "src_attn.linear_out.weight"], sd[key + "src_attn.linear_out.bias"] = raw[p + "cout.w"], raw[p + "cout.b"]
    sd["decoder.after_norm.weight"], sd["decoder.after_norm.bias"] = raw["dec_after_norm.g"], raw["dec_after_norm.b"]
    sd["decoder.output_layer.weight"], sd["decoder.output_layer.bias"] = raw["out.w"], raw["out.b"]
    torch.save(sd, folder / "model.pt")
    _mvn(folder / "am.mvn", raw["cmvn_means"], raw["cmvn_vars"])
    (folder / "config.yaml").write_text(f"encoder_conf:\n  attention_heads: {d.n_heads}\npredictor_conf:\n  tail_threshold: {d.tail_threshold}\n")
