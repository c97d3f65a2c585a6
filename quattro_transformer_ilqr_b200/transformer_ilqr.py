"""``TransformerILQR`` — load / save / predict wrapper around the CUDA transformer kernels.

Mirror of quattro_ilqr_tf/transformer_ilqr.py:24-325 of the reference: same constructor, ``load``
(:259-304), ``save`` (:213-254), ``predict`` (:311-325), ``fit`` (:102-208), attributes ``model``,
``normalizer``, ``device``, ``prompt_len``, ``target_len``.  ``predict`` and the in-solver
``forward_device`` run csrc/qx_tf_*.cu through the C ABI (``qx_tf_forward``): normalise -> embed ->
encoder layers -> output head -> de-normalise -> unpack, with no PyTorch operator on the path.

Checkpoints saved with ``quant_mode="float16"`` (both shipped ones) are loaded as the same fp16
values widened to fp32: the kernels compute in fp32, which is MORE accurate than the reference's
fp16 CPU inference (measured gap of the reference's own two modes: ~3e-4 relative).
``quant_mode="int8"`` checkpoints (torch dynamic quantisation of the nn.Linear modules) are read and
written in the reference's format; on load the int8 weights are de-quantised to fp32.
"""
from __future__ import annotations

import ctypes as C
import datetime
import os

import numpy as np
import torch

from . import _lib
from .transformer_model import DataNormalizer, TransformerPredictor


def _dequantize_dynamic_state_dict(sd):
    """state_dict of ``torch.quantization.quantize_dynamic(model, {nn.Linear}, qint8)`` (the reference's
    ``quant_mode="int8"`` checkpoints, transformer_ilqr.py:226-227 / :297-298) -> plain fp32 state_dict.
    Every quantised Linear stores ``<name>._packed_params._packed_params = (qint8 weight, bias)`` plus ``scale``,
    ``zero_point`` and ``dtype`` entries of its (unused, dynamic) output quantiser; the attention in_proj / out_proj and
    the LayerNorms are not nn.Linear modules and stay fp32.  The CUDA path then computes with the de-quantised int8
    weights and fp32 activations - the reference additionally quantises the activations of each Linear per call."""
    out = {}
    for key, val in sd.items():
        if key.endswith("._packed_params._packed_params"):
            prefix = key[: -len("._packed_params._packed_params")]
            qw, bias = val
            out[prefix + ".weight"] = qw.dequantize() if qw.is_quantized else qw
            if bias is not None:
                out[prefix + ".bias"] = bias
        elif key.endswith((".scale", ".zero_point", "._packed_params.dtype")):
            continue
        else:
            out[key] = val
    return out


class TransformerILQR:
    def __init__(self, state_dim, control_dim, prompt_len=10, d_model=64, nhead=8, num_decoder_layers=3,
                 dim_feedforward=128, dropout=0.1, max_seq_len=100, quant_mode="none"):
        self.state_dim, self.control_dim, self.prompt_len = state_dim, control_dim, prompt_len
        self.d_model, self.nhead, self.num_decoder_layers = d_model, nhead, num_decoder_layers
        self.dim_feedforward, self.dropout, self.max_seq_len = dim_feedforward, dropout, max_seq_len
        self.quant_mode = quant_mode
        self.target_len = None
        self.num_epochs = 0
        self.device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.model = None
        self.normalizer = DataNormalizer()
        self._packed = None          # (device, packed weights, TF32 hi/lo split of the dense weights)
        self.tensor_cores = True     # dense layers on tcgen05 (3xTF32); False = exact-fp32 CUDA cores
        self._shift = None
        self._solo = None            # 1-trajectory scratch engine state for predict()

    # ------------------------------------------------------------------ checkpoint wire format
    def _build_model(self):
        return TransformerPredictor(self.state_dim, self.control_dim, self.d_model, self.nhead,
                                    self.num_decoder_layers, self.dim_feedforward, self.dropout, self.max_seq_len,
                                    self.target_len, self.prompt_len)

    def load(self, model_path):
        # every field is a plain array or scalar: nothing in a checkpoint needs the pickle machinery
        data = np.load(os.path.join(model_path, "tf_model_normalizer.npz"), allow_pickle=False)
        nz = self.normalizer
        nz.x_mean, nz.x_std, nz.u_mean, nz.u_std = data["x_mean"], data["x_std"], data["u_mean"], data["u_std"]
        for key in ("target_len", "prompt_len", "state_dim", "control_dim", "d_model", "nhead",
                    "num_decoder_layers", "dim_feedforward", "max_seq_len", "num_epochs"):
            setattr(self, key, int(data[key]))
        self.dropout = float(data["dropout"])
        self.quant_mode = str(data["quant_mode"])
        self.model = self._build_model()
        sd = torch.load(os.path.join(model_path, "tf_model.pt"), map_location="cpu", weights_only=True)
        if self.quant_mode == "int8":
            sd = _dequantize_dynamic_state_dict(sd)
        self.model.load_state_dict({k: v.float() for k, v in sd.items()})
        self.model.eval()
        self._packed = None
        print(f"Model loaded from directory: {model_path}")
        return self

    def save(self, base_name):
        total = sum(p.numel() for p in self.model.parameters())
        size = f"{total / 1e6:.1f}M" if total >= 1e6 else (f"{total / 1e3:.1f}k" if total >= 1e3 else str(total))
        hyper = (f"decoder_dec{self.num_decoder_layers}_dmodel{self.d_model}_nhead{self.nhead}_ff{self.dim_feedforward}"
                 f"_drop{self.dropout}_epoch{self.num_epochs}_promptlen{self.prompt_len}")
        stamp = datetime.datetime.now().strftime("%Y%m%d_%H%M")
        dir_name = f"{stamp}_{base_name}_{hyper}_{size}"
        os.makedirs(dir_name, exist_ok=True)
        sd = self.model.state_dict()
        if self.quant_mode == "float16":
            sd = {k: v.half() for k, v in sd.items()}
        elif self.quant_mode == "int8":
            # the reference's wire format (transformer_ilqr.py:226-227): torch's dynamic quantisation of the nn.Linear
            # modules, applied to a CPU copy (the live fp32 module stays as it is)
            import copy
            q = torch.quantization.quantize_dynamic(copy.deepcopy(self.model).cpu().float(), {torch.nn.Linear},
                                                    dtype=torch.qint8)
            sd = q.state_dict()
        torch.save(sd, os.path.join(dir_name, "tf_model.pt"))
        nz = self.normalizer
        np.savez(os.path.join(dir_name, "tf_model_normalizer.npz"), x_mean=nz.x_mean, x_std=nz.x_std,
                 u_mean=nz.u_mean, u_std=nz.u_std, target_len=self.target_len, prompt_len=self.prompt_len,
                 state_dim=self.state_dim, control_dim=self.control_dim, d_model=self.d_model, nhead=self.nhead,
                 num_decoder_layers=self.num_decoder_layers, dim_feedforward=self.dim_feedforward,
                 dropout=self.dropout, max_seq_len=self.max_seq_len, num_epochs=self.num_epochs,
                 quant_mode=self.quant_mode)
        print(f"Model saved to {dir_name}")
        return dir_name

    # ------------------------------------------------------------------ device side
    def _m(self):
        return self.control_dim // (1 + self.state_dim)

    def config(self):
        cfg = _lib.QxTfConfig()
        cfg.state_dim, cfg.control_dim, cfg.d_model, cfg.nhead = self.state_dim, self.control_dim, self.d_model, self.nhead
        cfg.num_layers, cfg.dim_ff = self.num_decoder_layers, self.dim_feedforward
        cfg.target_len, cfg.prompt_len = self.target_len, self.prompt_len
        cfg.horizon = self.target_len + self.prompt_len
        cfg.m = self._m()
        return cfg

    def pack_weights(self):
        """Flatten normaliser + state_dict into the order of csrc/qx_tf_common.cuh (host tensor)."""
        sd = {k: v.detach().float().cpu() for k, v in self.model.state_dict().items()}
        nz = self.normalizer
        L = (self.target_len + self.prompt_len + 1) + self.prompt_len + self.target_len
        parts = [torch.as_tensor(np.asarray(a, dtype=np.float32)) for a in (nz.x_mean, nz.x_std, nz.u_mean, nz.u_std)]
        parts += [sd["state_embed.weight"], sd["state_embed.bias"], sd["control_embed.weight"], sd["control_embed.bias"],
                  sd["target_embedding"], sd["pos_encoder.pe"][0, :L]]
        for l in range(self.num_decoder_layers):
            p = f"transformer_decoder.layers.{l}."
            parts += [sd[p + k] for k in ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
                                          "self_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight",
                                          "linear2.bias", "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")]
        parts += [sd["output_linear.weight"], sd["output_linear.bias"]]
        def pad32(t):                      # every tensor starts on a 32-float boundary (qx_tf_common.cuh)
            t = t.reshape(-1)
            return torch.cat([t, t.new_zeros((-t.numel()) % 32)])
        flat = torch.cat([pad32(p) for p in parts]).contiguous()
        cfg = self.config()
        expect = _lib.load().qx_tf_weight_floats(C.byref(cfg))
        if flat.numel() != expect:
            raise RuntimeError(f"packed weights have {flat.numel()} floats, the library expects {expect}")
        return flat

    def _device_weights(self, device):
        """(packed weights, hi/lo split of the dense weights) resident on `device`."""
        if self._packed is None or self._packed[0] != device:
            lib = _lib.load()
            cfg = self.config()
            w = self.pack_weights().to(device)
            split = torch.empty(lib.qx_tf_split_floats(C.byref(cfg)), dtype=torch.float32, device=device)
            with torch.cuda.device(device):      # the launch needs `device` current, whatever the caller has set
                _lib.check(lib.qx_tf_prepare(C.byref(cfg), C.c_void_p(w.data_ptr()), C.c_void_p(split.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream(device).cuda_stream)), "qx_tf_prepare")
            self._packed = (device, w, split)
        return self._packed[1], self._packed[2]

    def invalidate(self):
        """Call after mutating ``model`` / ``normalizer`` in place (e.g. after training)."""
        self._packed = None

    def set_shift(self, shift, device):
        """x_seq_err = x_seq + shift, shift = state_offset - x_ref (quattro_ilqr_tf.py:504/:532)."""
        self._shift = torch.as_tensor(np.asarray(shift, dtype=np.float64), device=device)

    def graph_key(self, engine):
        """What a captured ``forward_device`` launch sequence depends on besides the engine's own buffers: the weight
        and split buffers, the state shift, the activations workspace and the kernel-path switches (read from the
        environment per call).  ``BatchedILQR.iterate_pair`` keys its CUDA graphs on it."""
        packed = self._packed
        shift = getattr(engine, "_tf_shift", None)
        if shift is None:
            shift = self._shift
        ws = getattr(engine, "_tf_ws", None)
        env = tuple(os.environ.get(k, "") for k in ("QX_TF_FUSE_LN", "QX_TF_NO_PRUNE", "QX_TF_GEMM_STREAM", "QX_TF_GEMM_VARIANT",
                                                     "QX_TF_ATTN_PAIR", "QX_TF_ATTN_MMA", "QX_TF_ATTN_DBG", "QX_TF_GEMM_DBG", "QX_TF_GEMM_F16",
                                                     "QX_TF_FFN_FUSE", "QX_TF_TAIL_FUSE", "QX_TF_ATTN_QK16", "QX_TF_UNPACK_TILE"))
        return (id(self), bool(self.tensor_cores), None if packed is None else (packed[1].data_ptr(), packed[2].data_ptr()),
                None if shift is None else shift.data_ptr(), None if ws is None else ws.data_ptr(), env)

    def forward_device(self, engine):
        """Predict gains for t < target_len from the engine's current trajectories and the Riccati
        tail already in ``engine.gains``; writes the predicted records into ``engine.gains``."""
        lib = _lib.load()
        cfg = self.config()
        dev = engine.device
        # The kernels index the trajectories and the gain records with the CHECKPOINT's horizon and dimensions
        # (a-Q(12): the reference silently mis-aligns the gains when they differ): an engine of another shape would be
        # read and written out of bounds.
        if (engine.N, engine.n, engine.m) != (cfg.horizon, self.state_dim, cfg.m):
            raise ValueError(f"engine shape (N, n, m) = {(engine.N, engine.n, engine.m)} does not match the checkpoint: "
                             f"horizon = target_len + prompt_len = {cfg.horizon}, state_dim = {self.state_dim}, "
                             f"m = control_dim / (1 + state_dim) = {cfg.m}")
        w, split = self._device_weights(dev)
        need = lib.qx_tf_workspace_bytes(C.byref(cfg), engine.B)
        # the activations live with the ENGINE, not with the wrapper: slices of a batch that run on their own streams
        # (pipeline.StreamedILQR) share the weights but must not share a workspace
        ws = getattr(engine, "_tf_ws", None)
        if ws is None or ws.device != dev or ws.numel() * 4 < need:
            ws = engine._tf_ws = torch.empty(need // 4, dtype=torch.float32, device=dev)
        # the state shift belongs to the engine (StreamedILQR gives every slice its own before any stream work starts);
        # the wrapper's own is the default for engines that never set one
        shift = getattr(engine, "_tf_shift", None)
        if shift is None:
            if self._shift is None or self._shift.device != dev:
                self.set_shift(np.zeros(self.state_dim), dev)
            shift = self._shift
        c = engine.cur
        with torch.cuda.device(dev):
            _lib.check(lib.qx_tf_forward(C.byref(cfg), C.c_void_p(w.data_ptr()),
                                         C.c_void_p(split.data_ptr() if self.tensor_cores else 0),
                                         C.c_void_p(engine.x[c].data_ptr()),
                                         C.c_void_p(shift.data_ptr()), C.c_void_p(engine.gains.data_ptr()),
                                         engine.model.model_id, C.c_void_p(engine.status.data_ptr()),
                                         C.c_void_p(ws.data_ptr()), engine.B, engine.stream), "qx_tf_forward")
        # embed, head, unpack + per layer QKV, attention, out-proj, add+LN, linear1, linear2, add+LN; on the tensor-core
        # path the two add+LN ride in the GEMM epilogues (QX_TF_FUSE_LN, qx_tf_simt.cu), except norm1 of the pruned last layer
        # (d = 128), and with the fp16 split linear1 + linear2 + norm2 are ONE launch (tf_ffn_kernel, QX_TF_FFN_FUSE)
        fuse = int(os.environ.get("QX_TF_FUSE_LN", "3")) if self.tensor_cores else 0
        pruned = self.tensor_cores and os.environ.get("QX_TF_NO_PRUNE", "0") != "1"
        f16 = (self.tensor_cores and os.environ.get("QX_TF_GEMM_F16", "1") != "0"
               and os.environ.get("QX_TF_GEMM_STREAM", "0") != "1"
               and self.d_model % 128 == 0 and self.dim_feedforward % 128 == 0)
        ffn = 1 if f16 and (fuse & 2) and os.environ.get("QX_TF_FFN_FUSE", "1") != "0" else 0
        # ... and out-proj + norm1 ride along when ff <= 256 (tf_tail_kernel, QX_TF_TAIL_FUSE), except in a pruned last layer
        tl = os.environ.get("QX_TF_TAIL_FUSE", "1")
        tail = ffn and (fuse & 1) and (tl == "2" or (tl == "1" and self.dim_feedforward <= 256))
        tail_layers = (self.num_decoder_layers - (1 if pruned else 0)) if tail else 0
        engine.launches += (3 + (7 - (fuse & 1) - ((fuse >> 1) & 1) - ffn) * self.num_decoder_layers - tail_layers
                            + (1 if (fuse & 1) and pruned else 0))

    def predict(self, x_seq, kK_seq):
        """x_seq (T, n) [or (B, T, n)], kK_seq (>= prompt_len, c): de-normalised predictions
        (target_len, c) as float32 — reference ``predict`` (transformer_ilqr.py:311-325)."""
        from .device_models import DeviceModel
        from .engine import BatchedILQR
        x = np.asarray(x_seq, dtype=np.float64)
        kK = np.asarray(kK_seq, dtype=np.float64)
        single = x.ndim == 2
        if single:
            x, kK = x[None], kK[None]
        B, N, n, m, P = x.shape[0], x.shape[1] - 1, self.state_dim, self._m(), self.prompt_len
        name = {(4, 1): "cartpole", (12, 4): "quadrotor"}.get((n, m))
        if name is None:
            raise ValueError("predict needs a built-in system shape (4x1 or 12x4)")
        if x.ndim != 3 or x.shape[2] != n or N != self.target_len + P:
            raise ValueError(f"x_seq must be (horizon + 1, state_dim) = ({self.target_len + P + 1}, {n}) "
                             f"[optionally batched], got {tuple(np.shape(x_seq))}")
        if kK.ndim != 3 or kK.shape[0] != B or kK.shape[1] < P or kK.shape[2] != m * (1 + n):
            raise ValueError(f"kK_seq must be (>= prompt_len, control_dim) = (>= {P}, {m * (1 + n)}) "
                             f"[optionally batched like x_seq], got {tuple(np.shape(kK_seq))}")
        solo = self._solo
        if solo is None or solo.B != B or solo.N != N or solo.model.name != name:
            solo = self._solo = BatchedILQR(DeviceModel(name), B, N)
        solo.x[solo.cur].copy_(torch.as_tensor(x).to(solo.device).permute(1, 2, 0))
        # prompt rows -> the last P gain records, inverse of the pack [k | K row-major]
        K_full = torch.zeros((B, N, m, n), dtype=torch.float32)
        k_full = torch.zeros((B, N, m), dtype=torch.float32)
        tail = torch.as_tensor(kK[:, -P:, :], dtype=torch.float32)
        k_full[:, N - P:] = tail[..., :m]
        K_full[:, N - P:] = tail[..., m:].reshape(B, P, m, n)
        solo.load_gains(K_full, k_full)
        solo.status.zero_()
        # predict() takes x_seq_err as is (transformer_ilqr.py:311-316): zero shift, on the scratch engine only - a
        # shift the caller installed for its own solver stays untouched
        if getattr(solo, "_tf_shift", None) is None:
            solo._tf_shift = torch.zeros(n, dtype=torch.float64, device=solo.device)
        self.forward_device(solo)
        K, k = solo.gains_batch()
        T = self.target_len
        pred = torch.cat([k[:, :T, :, None], K[:, :T]], dim=-1).reshape(B, T, m * (1 + n)).cpu().numpy()
        return pred[0] if single else pred

    # ------------------------------------------------------------------ training (PyTorch autograd; off the hot path)
    def _create_dataset(self, df):
        xs, kKs = [], []
        for _, row in df.iterrows():
            x = np.array(row["x_seq"], dtype=np.float32)
            k = np.array(row["k_seq"], dtype=np.float32)[..., None]
            K = np.array(row["K_seq"], dtype=np.float32)
            kK = np.concatenate([k, K], axis=-1).reshape(k.shape[0], -1)
            if kK.shape[0] <= self.prompt_len:
                continue
            xs.append(x)
            kKs.append(kK)
        return np.array(xs), np.array(kKs)

    def fit(self, df, test_df=None, num_epochs=50, batch_size=16, learning_rate=1e-3, patience=5):
        x_data, kK_data = self._create_dataset(df)
        T = x_data.shape[1]
        self.target_len = T - self.prompt_len
        nz = self.normalizer
        nz.fit(x_data, kK_data)
        xin = torch.tensor(nz.transform_x(x_data)).float()
        kK = torch.tensor(nz.transform_u(kK_data)).float()
        prompt, target = kK[:, -self.prompt_len:], kK[:, : T - self.prompt_len]
        self.model = self._build_model().to(self.device)
        self.num_epochs = num_epochs
        opt = torch.optim.Adam(self.model.parameters(), lr=learning_rate)
        loss_fn = torch.nn.MSELoss()
        if test_df is not None:
            xt, kKt = self._create_dataset(test_df)
            xt = torch.tensor(nz.transform_x(xt)).float().to(self.device)
            kKt = torch.tensor(nz.transform_u(kKt)).float().to(self.device)
        best, stale, best_state = float("inf"), 0, None
        self.train_loss_history, self.test_loss_history = [], []
        ds = torch.utils.data.TensorDataset(xin, prompt, target)
        loader = torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=True)
        for epoch in range(num_epochs):
            self.model.train()
            tot = 0.0
            for xb, pb, tb in loader:
                xb, pb, tb = xb.to(self.device), pb.to(self.device), tb.to(self.device)
                opt.zero_grad()
                loss = loss_fn(self.model(xb, pb), tb)
                loss.backward()
                opt.step()
                tot += loss.item() * xb.size(0)
            self.train_loss_history.append(tot / len(ds))
            if test_df is not None:
                self.model.eval()
                with torch.no_grad():
                    tl = loss_fn(self.model(xt, kKt[:, -self.prompt_len:]), kKt[:, : xt.shape[1] - self.prompt_len]).item()
                self.test_loss_history.append(tl)
                if tl < best:
                    best, stale = tl, 0
                    best_state = {k: v.clone() for k, v in self.model.state_dict().items()}
                else:
                    stale += 1
                if stale >= patience:
                    self.model.load_state_dict(best_state)
                    break
        self.model.eval()
        self._packed = None
        return self
