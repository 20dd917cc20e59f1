"""``Xtractor`` -- drop-in for ``sidekit.nnet.xvector.Xtractor`` on the inference path
(sidekit/nnet/xvector.py:419-686 constructor, :876-907 forward, :909-924 context_size).

Same constructor signature, same module tree and ``state_dict`` keys, same
return convention: ``(logits, embedding)`` for the margin losses, the bare embedding for
``loss='cce'`` with ``is_eval=True``.  The torch modules are parameter containers; ``forward`` hands
the whole pass to the native engine (csrc/engine.cu) through the C ABI.  Two deliberate deviations
from the reference as shipped, both required for its ``forward`` to run at all (SURVEY.md finding 4):
``halfresnet34`` pools with ``AttentivePooling(256, 10, global_context=True)`` (the shipped
``(256, 80)`` cannot consume the trunk output) and ``MfccFrontEnd.forward`` accepts ``is_eval``.
"""
import ctypes
import weakref
from collections import OrderedDict

import torch

from .. import _lib
from .loss import ArcMarginProduct, SoftmaxAngularProto
from .pooling import AttentivePooling, MeanStdPooling
from .preprocessor import MelSpecFrontEnd, MfccFrontEnd
from .res_net import PreHalfResNet34, PreResNet34, PreFastResNet34
from ..detplot import eer  # noqa: F401  (sidekit.nnet.xvector.eer, xvector.py:101-209)

_ARCHI_ID = {"halfresnet34": 0, "xvector": 1, "resnet34": 2, "fastresnet34": 3}
_RESNETS = ("halfresnet34", "resnet34", "fastresnet34")


class _NativeHandle:
    """Owns one ``skb_xtractor_t``; destroyed with the Python object."""

    def __init__(self, archi, state_dict, compute_dtype, margin_s):
        names, tensors = [], []
        for k, v in state_dict.items():
            if not v.is_floating_point():
                continue
            names.append(k.encode())
            tensors.append(v.detach().to("cpu", torch.float32).contiguous())
        n = len(names)
        c_names = (ctypes.c_char_p * n)(*names)
        c_data = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
        shape_arrays = [_lib.i64_array(list(t.shape) or [1]) for t in tensors]
        c_shapes = (_lib.c_i64_p * n)(*[ctypes.cast(a, _lib.c_i64_p) for a in shape_arrays])
        c_ndims = (ctypes.c_int * n)(*[max(t.dim(), 1) for t in tensors])
        out = ctypes.c_void_p()
        _lib.check(_lib.lib().skb_xtractor_create(_ARCHI_ID[archi], n, c_names, c_data, c_shapes, c_ndims,
                                                  compute_dtype, float(margin_s), ctypes.byref(out)))
        self.ptr = out
        self.overflow_seen = 0          # cumulative fp16-saturation count already reported (skb_xtractor_overflow_count)

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().skb_xtractor_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Xtractor(torch.nn.Module):
    """x-vector extractor; see the module docstring.  ``compute_dtype``: 'fp16' (default) or 'bf16'
    tensor-core operands, fp32 accumulation -- an extension, not a reference argument."""

    def __init__(self, speaker_number, model_archi="xvector", loss=None, norm_embedding=False, aam_margin=0.2,
                 aam_s=30, embedding_size=256, compute_dtype="fp16"):
        super().__init__()
        self.speaker_number = speaker_number
        self.feature_size = None
        self.norm_embedding = norm_embedding
        self.model_archi = model_archi
        self.compute_dtype = compute_dtype
        self._native = None
        self._native_device = None

        if model_archi == "xvector":
            self.input_nbdim = 2
            if loss not in ["cce", "aam"]:
                raise NotImplementedError("The valid loss are for now cce and aam ")
            self.loss = loss
            self.activation = torch.nn.LeakyReLU(0.2)
            self.preprocessor = MfccFrontEnd()
            self.feature_size = self.preprocessor.n_mfcc
            self.sequence_network = torch.nn.Sequential(OrderedDict([
                ("conv1", torch.nn.Conv1d(self.feature_size, 512, 5, dilation=1)),
                ("activation1", torch.nn.LeakyReLU(0.2)),
                ("batch_norm1", torch.nn.BatchNorm1d(512)),
                ("conv2", torch.nn.Conv1d(512, 512, 3, dilation=2)),
                ("activation2", torch.nn.LeakyReLU(0.2)),
                ("batch_norm2", torch.nn.BatchNorm1d(512)),
                ("conv3", torch.nn.Conv1d(512, 512, 3, dilation=3)),
                ("activation3", torch.nn.LeakyReLU(0.2)),
                ("batch_norm3", torch.nn.BatchNorm1d(512)),
                ("conv4", torch.nn.Conv1d(512, 512, 1)),
                ("activation4", torch.nn.LeakyReLU(0.2)),
                ("batch_norm4", torch.nn.BatchNorm1d(512)),
                ("conv5", torch.nn.Conv1d(512, 1536, 1)),
                ("activation5", torch.nn.LeakyReLU(0.2)),
                ("batch_norm5", torch.nn.BatchNorm1d(1536)),
            ]))
            self.embedding_size = embedding_size
            self.stat_pooling = MeanStdPooling()
            self.before_speaker_embedding = torch.nn.Sequential(OrderedDict([
                ("linear6", torch.nn.Linear(3072, self.embedding_size))]))
            if self.loss == "aam":
                self.after_speaker_embedding = ArcMarginProduct(self.embedding_size, int(self.speaker_number),
                                                                s=64, m=0.2, easy_margin=False)
                self._margin_s = 64.0
            else:   # 'cce': the classifier stack only matters in training; forward(is_eval=True) returns x
                self.after_speaker_embedding = torch.nn.Sequential(OrderedDict([
                    ("activation6", torch.nn.LeakyReLU(0.2)),
                    ("batch_norm6", torch.nn.BatchNorm1d(512)),
                    ("dropout6", torch.nn.Dropout(p=0.05)),
                    ("linear7", torch.nn.Linear(512, 512)),
                    ("activation7", torch.nn.LeakyReLU(0.2)),
                    ("batch_norm7", torch.nn.BatchNorm1d(512)),
                    ("linear8", torch.nn.Linear(512, int(self.speaker_number)))]))
                self._margin_s = 0.0
        elif model_archi == "halfresnet34":
            self.preprocessor = MelSpecFrontEnd(n_fft=1024, win_length=400, hop_length=160, n_mels=80)
            self.sequence_network = PreHalfResNet34()
            self.embedding_size = embedding_size
            self.before_speaker_embedding = torch.nn.Sequential(OrderedDict([
                ("lin_be", torch.nn.Linear(in_features=5120, out_features=self.embedding_size, bias=False)),
                ("bn_be", torch.nn.BatchNorm1d(self.embedding_size))]))
            self.stat_pooling = AttentivePooling(256, 10, global_context=True)
            self.loss = loss
            # xvector.py:585-593: 'aam' -> ArcMarginProduct, 'aps' -> SoftmaxAngularProto, anything else (None, 'cce') builds NO
            # after_speaker_embedding; forward then returns the bare embedding (:896-907)
            self._margin_s = 0.0
            if self.loss == "aam":
                self.after_speaker_embedding = ArcMarginProduct(self.embedding_size, int(self.speaker_number),
                                                                s=30, m=0.2, easy_margin=False)
                self._margin_s = 30.0
            elif self.loss == "aps":
                self.after_speaker_embedding = SoftmaxAngularProto(int(self.speaker_number), emb_dim=self.embedding_size)
        elif model_archi == "resnet34":
            # xvector.py:516-540.  Same deviation as halfresnet34: the shipped AttentivePooling(256, 80, ...) cannot consume
            # the trunk's 256 x 10 output; the 5120-wide Linear that follows shows the intended pooling.
            self.preprocessor = MelSpecFrontEnd(n_fft=1024, win_length=400, hop_length=160, n_mels=80)
            self.sequence_network = PreResNet34()
            self.embedding_size = embedding_size
            self.before_speaker_embedding = torch.nn.Linear(in_features=5120, out_features=self.embedding_size)
            self.stat_pooling = AttentivePooling(256, 10, global_context=True)
            self.loss = "aam"
            self.after_speaker_embedding = ArcMarginProduct(self.embedding_size, int(self.speaker_number), s=30.0, m=0.20,
                                                            easy_margin=False)
            self._margin_s = 30.0
        elif model_archi == "fastresnet34":
            # xvector.py:539-567.  The shipped AttentivePooling(128, 80, global_context=False) expects 10240 inputs while the
            # trunk emits 128 x 10 and the Linear that follows has 2560 = 2 * 1280 inputs: num_freqs = 10 is what runs.
            self.preprocessor = MelSpecFrontEnd()
            self.sequence_network = PreFastResNet34()
            self.embedding_size = embedding_size
            self.before_speaker_embedding = torch.nn.Linear(in_features=2560, out_features=self.embedding_size)
            self.stat_pooling = AttentivePooling(128, 10, global_context=False)
            self.loss = loss
            if self.loss == "aam":
                self.after_speaker_embedding = ArcMarginProduct(self.embedding_size, int(self.speaker_number), s=30, m=0.2,
                                                                easy_margin=False)
                self._margin_s = 30.0
            else:
                raise NotImplementedError("only loss='aam' is implemented for fastresnet34 (inference hot path)")
        else:
            raise NotImplementedError("model_archi %r: the B200 engine implements 'halfresnet34', 'resnet34', 'fastresnet34' "
                                      "and 'xvector'" % (model_archi,))
        self.preprocessor.__dict__["_owner"] = weakref.ref(self)

    # ------------------------------------------------------------------ native engine management
    def _apply(self, fn, *a, **k):
        self._native = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._native = None
        return super().load_state_dict(*a, **k)

    def refresh(self):
        """Re-pack the weights after an in-place parameter edit."""
        self._native = None

    def _handle(self, device):
        if not torch.cuda.is_available():
            raise RuntimeError("sidekit_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if self._native is None or self._native_device != device:
            with torch.cuda.device(device):
                dt = {"fp16": 0, "bf16": 1}[self.compute_dtype]
                self._native = _NativeHandle(self.model_archi, self.state_dict(), dt, self._margin_s)
            self._native_device = device
        return self._native.ptr

    # ------------------------------------------------------------------ forward
    def _pack(self, waves):
        lengths = [int(w.shape[-1]) for w in waves]
        flat = torch.cat([w.reshape(-1) for w in waves]) if len(waves) > 1 else waves[0].reshape(-1)
        return flat.contiguous().float(), lengths

    def reserve(self, max_utts, max_audio_seconds, device=None, sample_rate=16000):
        """Extension for bulk extraction: allocate every work buffer and geometry-plan slot for packed batches of up to
        ``max_utts`` utterances / ``max_audio_seconds`` of audio now (``skb_xtractor_reserve``), so the run itself never
        allocates."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        h = self._handle(device)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().skb_xtractor_reserve(h, int(max_utts), int(max_audio_seconds * sample_rate), _lib.stream_ptr()))

    def check_overflow(self):
        """fp16 range guard: raises ``OverflowError`` when an activation stored since the last check saturated the fp16 range
        (|x| >= 65504; the kernels clamp instead of producing inf, so the embeddings of those calls are wrong, not NaN).
        Synchronises the current stream.  ``forward`` / ``extract_varlen`` / ``extract_stream`` call it themselves;
        ``extract_packed`` does not (bulk loops check once at the end)."""
        nat = self._native
        if nat is None or self.compute_dtype != "fp16":
            return
        count = ctypes.c_int64(0)
        with torch.cuda.device(self._native_device):
            _lib.check(_lib.lib().skb_xtractor_overflow_count(nat.ptr, _lib.stream_ptr(), ctypes.byref(count)))
        if count.value > nat.overflow_seen:
            n = count.value - nat.overflow_seen
            nat.overflow_seen = count.value
            raise OverflowError("sidekit_b200: %d kernel threads stored activations beyond the fp16 range (65504); these weights "
                                "need Xtractor(..., compute_dtype='bf16')" % n)

    def _run(self, flat, lengths, norm_embedding, want_logits=True, emb_out=None):
        B = len(lengths)
        if self.loss not in ("aam", "aps"):
            want_logits = False
        on_cpu = not flat.is_cuda
        if not torch.cuda.is_available():
            raise RuntimeError("sidekit_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device()) if on_cpu else flat.device
        h = self._handle(device)
        lens = _lib.i64_array(lengths)
        E, S = self.embedding_size, int(self.speaker_number)
        with torch.cuda.device(device):
            if on_cpu:
                emb = torch.empty((B, E), dtype=torch.float32)
                logits = torch.empty((B, S), dtype=torch.float32) if want_logits else None
                _lib.check(_lib.lib().skb_xtractor_forward_host(h, flat.data_ptr(), lens, B, int(norm_embedding),
                                                               emb.data_ptr(), logits.data_ptr() if want_logits else None,
                                                               _lib.stream_ptr()))
            else:
                emb = emb_out if emb_out is not None else torch.empty((B, E), dtype=torch.float32, device=device)
                assert emb.is_contiguous() and tuple(emb.shape) == (B, E) and emb.dtype == torch.float32 and emb.device == device
                logits = torch.empty((B, S), dtype=torch.float32, device=device) if want_logits else None
                _lib.check(_lib.lib().skb_xtractor_forward(h, flat.data_ptr(), lens, B, int(norm_embedding),
                                                          emb.data_ptr(), logits.data_ptr() if want_logits else None,
                                                          _lib.stream_ptr()))
        return logits, emb

    def forward(self, x, is_eval=False, target=None, norm_embedding=True):
        """Same contract as the reference's ``forward`` for ``is_eval=True, target=None``:
        ``x`` is a (L,) or (B, L) float waveform at 16 kHz; returns ``(s*cos logits, F.normalize(embedding))``."""
        if not is_eval or target is not None:
            raise NotImplementedError("sidekit_b200 implements the inference path: call forward(x, is_eval=True)")
        if x.dim() == 1:
            x = x.unsqueeze(0)
        assert x.dim() == 2, "expected a (B, L) or (L,) waveform tensor"
        B, L = x.shape
        flat = x.contiguous().float().reshape(-1)
        logits, emb = self._run(flat, [L] * B, norm_embedding)
        self.check_overflow()
        if self.loss not in ("aam", "aps"):   # xvector.py:896-907: 'cce' at eval time, or no loss: the bare (l2-normalised or raw) embedding
            return self._pre_embedding(B, flat)
        return logits, emb

    def _pre_embedding(self, B, like):
        """x before the final F.normalize (what loss='cce' returns at eval time)."""
        on_cpu = not like.is_cuda
        device = self._native_device
        out = torch.empty((B, self.embedding_size), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().skb_xtractor_pre_embedding(self._native.ptr, B, out.data_ptr(), _lib.stream_ptr()))
        return out.cpu() if on_cpu else out

    def extract_varlen(self, waves, norm_embedding=True, want_logits=False):
        """Extension: one call for a list of utterances of DIFFERENT lengths, packed without padding
        (every reduction in the engine is per utterance, so results equal one-by-one extraction)."""
        flat, lengths = self._pack(list(waves))
        logits, emb = self._run(flat, lengths, norm_embedding, want_logits)
        self.check_overflow()
        return (logits, emb) if want_logits else emb

    def extract_packed(self, flat, lengths, norm_embedding=True, want_logits=False, out=None):
        """Extension for bulk extraction: ``flat`` already holds the utterances back to back (1-D fp32, CUDA or
        -- ideally pinned -- host memory), ``lengths`` their sample counts.  Host input goes through
        ``skb_xtractor_forward_host`` (H2D, forward, D2H inside one native call).  ``out``: a (n, E) CUDA tensor (e.g. a
        slice of the shard's embedding block) that receives the embeddings instead of a fresh allocation.
        Asynchronous: the fp16 range guard is NOT checked here, call ``check_overflow()`` after the loop."""
        logits, emb = self._run(flat, [int(v) for v in lengths], norm_embedding, want_logits, emb_out=out if flat.is_cuda else None)
        return (logits, emb) if want_logits else emb

    def extract_stream(self, batches, norm_embedding=True, device_out=None):
        """Extension for bulk extraction from HOST memory with copy / compute overlap: ``batches`` is a sequence of
        ``(flat, lengths)`` pairs as for ``extract_packed`` (``flat`` ideally pinned).  The waveforms of batch i+1 travel
        over PCIe on a second stream while batch i is being embedded (two device staging buffers, events both ways);
        the embeddings come back into one pinned host tensor.  Returns a list of (n_i, E) host tensors (views).
        ``device_out``: optional (sum n_i, E) CUDA tensor that also keeps the embeddings on the device, in batch order
        (what the multi-GPU all-gather consumes)."""
        if not torch.cuda.is_available():
            raise RuntimeError("sidekit_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        batches = list(batches)
        device = torch.device("cuda", torch.cuda.current_device())
        compute = torch.cuda.current_stream(device)
        if getattr(self, "_copy_stream", None) is None or self._copy_stream.device != device:
            self._copy_stream = torch.cuda.Stream(device)
            self._stage = [None, None]
            self._out_host = None
        copy_s = self._copy_stream
        total = sum(len(l) for _, l in batches)
        if self._out_host is None or self._out_host.shape[0] < total:
            self._out_host = torch.empty((total, self.embedding_size), dtype=torch.float32, pin_memory=True)
        free_ev = [None, None]
        outs, row = [], 0
        # both staging buffers are sized for the largest batch of the call up front: growing one in the middle of the
        # pipeline costs a drain of both streams
        n_max = max((flat.numel() for flat, _ in batches), default=0)
        for k in range(2):
            if self._stage[k] is None or self._stage[k].numel() < n_max:
                compute.synchronize()                              # (re)allocation: nothing in flight may still use the old buffer
                copy_s.synchronize()
                self._stage[k] = torch.empty((n_max + n_max // 8,), dtype=torch.float32, device=device)
        for i, (flat, lengths) in enumerate(batches):
            k, n = i % 2, flat.numel()
            if self._stage[k] is None or self._stage[k].numel() < n:
                compute.synchronize()
                copy_s.synchronize()
                self._stage[k] = torch.empty((n + n // 8,), dtype=torch.float32, device=device)
            with torch.cuda.stream(copy_s):
                if free_ev[k] is not None:
                    copy_s.wait_event(free_ev[k])                  # the forward that read this staging buffer is done
                self._stage[k][:n].copy_(flat.reshape(-1), non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_s)
            compute.wait_event(ready)
            _, emb = self._run(self._stage[k][:n], [int(v) for v in lengths], norm_embedding, want_logits=False,
                               emb_out=None if device_out is None else device_out[row:row + len(lengths)])
            free_ev[k] = torch.cuda.Event()
            free_ev[k].record(compute)
            # the next waveform copy and this forward's geometry tables share the host-to-device copy engine and become
            # eligible at the same instant (when the previous forward ends): the tables must go first, or the forward
            # stalls for the whole 1.2 ms of the waveform copy (profiles/r02u_e2e_timeline.txt)
            _lib.check(_lib.lib().skb_xtractor_wait_tables(self._handle(device), ctypes.c_void_p(copy_s.cuda_stream)))
            dst = self._out_host[row:row + emb.shape[0]]
            dst.copy_(emb, non_blocking=True)
            outs.append(dst)
            row += emb.shape[0]
        compute.synchronize()
        self.check_overflow()
        return outs

    def _frontend(self, x):
        B, L = x.shape
        if not x.is_cuda:
            raise RuntimeError("sidekit_b200 has no CPU path: move the waveform to a CUDA device")
        h = self._handle(x.device)
        T = _lib.lib().skb_xtractor_num_frames(h, L)
        out = torch.empty((B, self.preprocessor.n_mfcc if self.model_archi == "xvector" else self.preprocessor.n_mels, T),
                          dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().skb_xtractor_frontend(h, x.contiguous().float().data_ptr(), _lib.i64_array([L] * B), B, T,
                                                        out.data_ptr(), _lib.stream_ptr()))
        return out

    def debug_stage(self, waves, stage):
        """Test hook: activation after ``stage`` ('stem', 'layer1.0' ... 'layer4.2', 'tdnn1'..'tdnn5', 'pooled')."""
        flat, lengths = self._pack(list(waves))
        h = self._handle(flat.device)
        B = len(lengths)
        T = [_lib.lib().skb_xtractor_num_frames(h, L) for L in lengths]
        if stage == "pooled":
            numel, shape = None, None
        per = ctypes.c_int64(0)
        if self.model_archi in _RESNETS:
            halves = (False, True, True, True)
            Ws = (80, 40, 20, 10)
            if self.model_archi == "halfresnet34":
                li = 0 if stage == "stem" else int(stage[5]) - 1 if stage.startswith("layer") else 3
                C = (32, 64, 128, 256)[li]
            elif self.model_archi == "fastresnet34":               # level 0 is carried with 32 channels (16 real + 16 zero)
                li = 0 if stage == "stem" else int(stage[5]) - 1 if stage.startswith("layer") else 3
                C = (32, 32, 64, 128)[li]
                Ws, halves = (40, 20, 10, 10), (False, True, True, False)
            else:                                                  # resolution level of layer1..layer7 (strides 1,2,1,2,1,2,1)
                li = 0 if stage == "stem" else (0, 1, 1, 2, 2, 3, 3)[int(stage[5]) - 1] if stage.startswith("layer") else 3
                C = (128, 128, 256, 256)[li]
            W = Ws[li]
            Hm = max(T)
            for l in range(1, li + 1):
                if halves[l]:
                    Hm = (Hm - 1) // 2 + 1
        else:
            C, W, Hm = (1536 if stage in ("tdnn5",) else 512), 1, max(T)
        size = B * {"halfresnet34": 5120, "resnet34": 5120, "fastresnet34": 2560, "xvector": 3072}[self.model_archi] \
            if stage == "pooled" else B * C * Hm * W
        out = torch.zeros(size, dtype=torch.float32, device=flat.device)
        with torch.cuda.device(flat.device):
            _lib.check(_lib.lib().skb_xtractor_debug_stage(h, flat.data_ptr(), _lib.i64_array(lengths), B, stage.encode(), Hm,
                                                           out.data_ptr(), ctypes.byref(per), _lib.stream_ptr()))
        return out.view(B, -1) if stage == "pooled" else out.view(B, C, Hm, W)

    def context_size(self):
        context = 1
        for name, module in self.sequence_network.named_modules():
            if name.startswith("conv"):
                context += module.dilation[0] * (module.kernel_size[0] - 1)
        return context
