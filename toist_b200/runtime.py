"""Glue between torch.autograd / nn.Module parameters and the kernel-level blocks.

* `ShadowBank` keeps a bf16 copy of every GEMM / convolution weight (FrozenBatchNorm scale folded per output
  channel, conv weights re-laid out OIHW -> OHWI) and refreshes all of them with one `toist_weight_prep` launch.
* Five autograd stages wrap the parts of the hot path; their backward passes are hand written on top of
  blocks.py, so torch only moves gradients between stages and accumulates them into `Parameter.grad` (which keeps
  DistributedDataParallel's reducer hooks working, reference main.py:336).

  BACKBONE  images -> NHWC features            (models/backbone.py:74-80, torchvision resnet)
  TEXT      token ids -> resized text features (models/transformer.py:129-138,487-492)
  ENCODER   features + text -> img_memory      (models/mdetr.py:383 input_proj, models/transformer.py:144-152)
  DECODER   img_memory -> hs                   (models/transformer.py:170-188)
  HEADS     hs -> logits / boxes / projections (models/mdetr.py:420-433)

  Each is a `Spec` (forward / backward launch sequence) run by the generic `StageFn`; with CUDA graphs enabled the
  sequences are captured once per shape signature and replayed (`GraphCache`).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

import torch

from . import blocks as Bk
from . import kernels as K
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID

BF = torch.bfloat16


# ------------------------------------------------------------------------------------------------ prefix views
class WView:
    """Read-only view of the flat weight dict under a name prefix."""

    __slots__ = ("d", "p")

    def __init__(self, d, p: str):
        self.d, self.p = d, p

    def __getitem__(self, k: str):
        return self.d[self.p + k]

    def sub(self, p: str) -> "WView":
        return WView(self.d, self.p + p)


class GView:
    """Gradient sink under a name prefix (dict-like: in / get / set)."""

    __slots__ = ("d", "p")

    def __init__(self, d, p: str):
        self.d, self.p = d, p

    def __contains__(self, k: str) -> bool:
        return (self.p + k) in self.d

    def __getitem__(self, k: str):
        return self.d[self.p + k]

    def __setitem__(self, k: str, v) -> None:
        self.d[self.p + k] = v


class RView:
    __slots__ = ("s", "p")

    def __init__(self, s: Set[str], p: str):
        self.s, self.p = s, p

    def __contains__(self, k: str) -> bool:
        return (self.p + k) in self.s


# ------------------------------------------------------------------------------------------------ saved-tensor packing
def _pack(obj, flat: List[torch.Tensor]):
    if isinstance(obj, torch.Tensor):
        flat.append(obj)
        return ("t", len(flat) - 1)
    if isinstance(obj, (tuple, list)):
        return ("l", [_pack(o, flat) for o in obj])
    return ("c", obj)


def _unpack(spec, flat: Sequence[torch.Tensor]):
    kind, val = spec
    if kind == "t":
        return flat[val]
    if kind == "l":
        return tuple(_unpack(s, flat) for s in val)
    return val


def _save(ctx, obj) -> None:
    flat: List[torch.Tensor] = []
    ctx.pack_spec = _pack(obj, flat)
    ctx.save_for_backward(*flat)


def _load(ctx):
    return _unpack(ctx.pack_spec, ctx.saved_tensors)


# ------------------------------------------------------------------------------------------------ shadow weights
_SHADOW_ALWAYS = os.environ.get("TOIST_SHADOW_ALWAYS", "0") != "0"
_CROSS_KV_HOIST = os.environ.get("TOIST_CROSS_KV_HOIST", "1") != "0"
_STEM_NHWC8 = os.environ.get("TOIST_STEM_NHWC8", "1") != "0"
_EMBEDDING_KEYS = ("word_embeddings", "position_embeddings", "token_type_embeddings", "query_embed")
RESNET_BLOCKS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}
STEM_LDK = 192  # 7*7*3 = 147 patch columns padded to a multiple of 64


class ShadowBank:
    """bf16 working copies of a model's weights + derived FrozenBatchNorm scale / shift vectors."""

    def __init__(self):
        self.w: Dict[str, torch.Tensor] = {}
        self.prep: Optional[K.WeightPrep] = None       # trunk convolutions
        self.prep_rest: Optional[K.WeightPrep] = None  # every other shadow
        self._tracked: List[torch.Tensor] = []
        self._ver = -1
        self._stale_rest = True
        self._sig = None

    def __deepcopy__(self, memo):  # EMA deep-copies the model (util/optim.py, main.py:322); rebuild lazily there
        return ShadowBank()

    @staticmethod
    def _signature(params, buffers) -> tuple:
        ps = tuple(p.data_ptr() for p in params.values())
        bs = tuple((b.data_ptr(), b._version) for b in buffers)
        return ps + bs

    def ensure(self, model, backbone_prefix: Optional[str], body, dirty: bool = True, only=None,
               defer_rest: bool = False) -> None:
        """(Re)builds the bank when parameters moved (device change, load of a new module) or BN buffers changed,
        then refreshes every shadow from its fp32 master.  `dirty=False` (nothing called Module._apply or
        load_state_dict since the last check) skips the pointer / version scan of ~1000 tensors.
        The refresh is two launches: the trunk's convolution weights (needed first, 0.26 GB of traffic) and everything
        else (RoBERTa, transformer, heads: 0.85 GB).  With `defer_rest` the caller issues the second one itself
        (`run_rest()`, on the text branch's stream, next to the trunk) instead of in front of it."""
        rebuilt = False
        if dirty or self._sig is None:
            params = {n: p for n, p in model.named_parameters() if only is None or n.startswith(only)}
            sig = self._signature(params, list(model.buffers()) if body is not None else [])
            if sig != self._sig:
                self._build(params, backbone_prefix, body)
                self._sig = sig
                self._tracked = list(params.values()) + (list(model.buffers()) if body is not None else [])
                rebuilt = True
        # The shadows are a cache of the fp32 masters: refresh them only when a master changed.  Every in-place update
        # that goes through torch (torch.optim.*, load_state_dict, copy_ on state_dict() entries) bumps the tensor's
        # version counter; toist_b200.util.optim.FusedAdamW bumps it explicitly.  (Writes through `.data` do not:
        # TOIST_SHADOW_ALWAYS=1 restores the unconditional refresh.)  Cost: one pass over ~500 counters, ~40 us.
        ver = sum(t._version for t in self._tracked)
        self._stale_rest = rebuilt or _SHADOW_ALWAYS or ver != self._ver
        self._ver = ver
        if self._stale_rest:
            self.prep.run()
            self._refresh_stem7()
            if not defer_rest:
                self.run_rest()

    def _refresh_stem7(self) -> None:
        st = getattr(self, "_stem", None)
        if st is None:
            return
        conv_w, scale, w7 = st
        if self._stem_ver == conv_w._version:  # the stem is frozen in every recipe: this runs once per (re)build / load
            return
        with torch.no_grad():
            w7.view(64, 7, 8, 8)[:, :, :7, :3].copy_((conv_w.detach() * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1))
        self._stem_ver = conv_w._version

    def run_rest(self) -> None:
        if self._stale_rest:
            self.prep_rest.run()
            for dst, parts in getattr(self, "_cat", ()):
                torch.cat(parts, out=dst)
            self._stale_rest = False

    def _build(self, params, backbone_prefix: Optional[str], body) -> None:
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise RuntimeError("toist_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        self._stem = None
        trunk = K.WeightPrep(dev)   # conv + FrozenBatchNorm pairs of the backbone
        prep = K.WeightPrep(dev)    # everything else
        w: Dict[str, torch.Tensor] = {}
        conv_names = set()
        # ---- backbone: conv + FrozenBatchNorm pairs
        for conv_name, bn_name in (body.conv_bn_pairs() if body is not None else ()):
            conv_w = params[backbone_prefix + conv_name + ".weight"]
            bn = body.get_submodule(bn_name)
            scale = (bn.weight * (bn.running_var + 1e-5).rsqrt()).float().contiguous()  # models/backbone.py:54-57
            shift = (bn.bias - bn.running_mean * scale).float().contiguous()
            cout, cin, kh, kw = conv_w.shape
            if conv_name == "conv1":  # 7x7 stem, consumed as an im2col GEMM
                sh = torch.zeros((cout, STEM_LDK), dtype=BF, device=dev)
                trunk.add(conv_w.detach(), sh, cout, cin * kh * kw, STEM_LDK, scale, taps=kh * kw)
                if (cout, cin, kh, kw) == (64, 3, 7, 7) and _STEM_NHWC8:
                    # the same filter as [64, 7 rows, 8 kernel columns x 8 channels] for kernels.stem_conv7x7 (kernel column 7
                    # and channels 3..7 are zero); derived from the master in _refresh_stem7 when it changed
                    w[backbone_prefix + "conv1.weight7"] = torch.zeros((64, 7 * 64), dtype=BF, device=dev)
                    self._stem = (conv_w, scale, w[backbone_prefix + "conv1.weight7"])
                    self._stem_ver = None
            else:
                sh = torch.empty((cout, kh, kw, cin), dtype=BF, device=dev)
                trunk.add(conv_w.detach(), sh, cout, cin * kh * kw, None, scale, taps=kh * kw)
            w[backbone_prefix + conv_name + ".weight"] = sh
            w[backbone_prefix + bn_name + ".scale"] = scale
            w[backbone_prefix + bn_name + ".shift"] = shift
            conv_names.add(backbone_prefix + conv_name + ".weight")
        # ---- everything else
        for name, p in params.items():
            if name in conv_names:
                continue
            qkv_part = ".attention.self." in name and name.endswith(("query.weight", "key.weight", "value.weight"))
            if p.dim() == 4 and p.shape[2] * p.shape[3] > 1:  # biased k x k convolution (mask head): OIHW -> OHWI
                cout, cin, kh, kw = p.shape
                sh = torch.empty((cout, kh, kw, cin), dtype=BF, device=dev)
                prep.add(p.detach(), sh, cout, cin * kh * kw, None, None, taps=kh * kw)
                w[name] = sh
            elif p.dim() >= 2 and not qkv_part and not any(k in name for k in _EMBEDDING_KEYS):
                rows = p.shape[0]
                cols = p.numel() // rows
                sh = torch.empty((rows, cols), dtype=BF, device=dev)
                prep.add(p.detach(), sh, rows, cols)
                w[name] = sh
            else:
                w[name] = p.detach()
        # ---- RoBERTa: q | k | v stacked so that the data gradient is one GEMM
        for name in list(params):
            if name.endswith("attention.self.query.weight"):
                base = name[: -len("query.weight")]
                E = params[name].shape[0]
                qkv = torch.empty((3 * E, E), dtype=BF, device=dev)
                for i, nm in enumerate(("query", "key", "value")):
                    prep.add(params[base + nm + ".weight"].detach(), qkv[i * E:(i + 1) * E], E, E)
                w[base + "qkv"] = qkv
        # ---- decoder: the key / value projections of the encoder memory of ALL layers as one [layers * E, E] weight each
        # (they depend on the memory only, not on the decoder state: one M = S * B, N = layers * E GEMM in front of the
        # decoder instead of two small GEMMs inside every layer)
        self._cat = []  # (destination fp32, [source slices]) refreshed by torch.cat in run_rest
        cross = sorted((n for n in params if n.endswith("cross_attn_image.in_proj_weight") and ".decoder.layers." in n),
                       key=lambda n: int(n.split(".layers.")[1].split(".")[0]))
        if cross and _CROSS_KV_HOIST:
            pre = cross[0].split("layers.")[0]  # "transformer.decoder."
            E = params[cross[0]].shape[1]
            L = len(cross)
            k_all = torch.empty((L * E, E), dtype=BF, device=dev)
            v_all = torch.empty((L * E, E), dtype=BF, device=dev)
            bk, bv = [], []
            for l, n in enumerate(cross):
                pw = params[n].detach()
                prep.add(pw[E: 2 * E], k_all[l * E:(l + 1) * E], E, E)
                prep.add(pw[2 * E:], v_all[l * E:(l + 1) * E], E, E)
                pb = params[n[: -len("weight")] + "bias"].detach()
                bk.append(pb[E: 2 * E])
                bv.append(pb[2 * E:])
            w[pre + "cross_kv.k_weight"], w[pre + "cross_kv.v_weight"] = k_all, v_all
            w[pre + "cross_kv.k_bias"] = torch.empty((L * E,), dtype=torch.float32, device=dev)
            w[pre + "cross_kv.v_bias"] = torch.empty((L * E,), dtype=torch.float32, device=dev)
            self._cat = [(w[pre + "cross_kv.k_bias"], bk), (w[pre + "cross_kv.v_bias"], bv)]
        self.w, self.prep, self.prep_rest = w, trunk, prep


def requires(names: Iterable[str], params: Sequence[torch.Tensor], enabled: bool) -> Set[str]:
    return {n for n, p in zip(names, params) if enabled and p.requires_grad}


def _grads_for(names: Sequence[str], g: Dict[str, torch.Tensor], params_shapes) -> tuple:
    out = []
    for n, shp in zip(names, params_shapes):
        t = g.get(n)
        out.append(None if t is None else t.view(shp))
    return tuple(out)


_EMPTY: frozenset = frozenset()


class Stage:
    """Static description of one autograd stage: which parameters it owns (ordered) and configuration."""

    def __init__(self, model, name: str, names: Sequence[str], **cfg):
        lookup = dict(model.named_parameters())
        self.name = name
        self.names = list(names)
        self.params = [lookup[n] for n in self.names]
        self.shapes = [tuple(p.shape) for p in self.params]
        self.__dict__.update(cfg)

    @classmethod
    def empty(cls, name: str, **cfg) -> "Stage":
        """A stage without parameters (the criterion)."""
        st = cls.__new__(cls)
        st.name, st.names, st.params, st.shapes = name, [], [], []
        st.__dict__.update(cfg)
        return st

    def req(self) -> Set[str]:
        """Names of this stage's parameters that want a gradient (empty under no_grad).  Cached: walking ~1000
        parameters per step costs more than several graph replays; `invalidate()` (ModelRuntime.refresh: after
        Module._apply / load_state_dict, and every 64 steps) rescans."""
        if not torch.is_grad_enabled():
            return _EMPTY
        r = self.__dict__.get("_req")
        if r is None:
            r = self.__dict__["_req"] = frozenset(requires(self.names, self.params, True))
        return r

    def invalidate(self) -> None:
        self.__dict__.pop("_req", None)
        self.__dict__.pop("_numel", None)

    def arena_numel(self, req) -> int:
        """fp32 elements of the gradient arena of one backward pass: every wanted gradient, 256-byte aligned."""
        cache = self.__dict__.setdefault("_numel", {})
        n = cache.get(req)
        if n is None:
            n = cache[req] = sum((_numel(shp) + 63) // 64 * 64 for nm, shp in zip(self.names, self.shapes) if nm in req)
        return n


class Call:
    """Per-invocation context handed to a Function (non-tensor argument)."""

    def __init__(self, stage: Stage, w: Dict[str, torch.Tensor], save: bool, graphs: "Optional[GraphCache]" = None,
                 drop_p: float = 0.0, seed: Optional[torch.Tensor] = None, **kw):
        self.stage, self.w = stage, w
        self.req = stage.req() if save else _EMPTY
        self.save = save  # keep activations for a backward pass (some tensor upstream or here wants a gradient)
        self.graphs = graphs
        self.drop_p = float(drop_p) if seed is not None else 0.0  # training-mode dropout probability of this stage
        self.seed = seed  # device int64 [1]; rewritten by the host every step, read by the kernels
        self.extra = tuple(sorted(kw.items()))
        self.__dict__.update(kw)

    def signature(self) -> tuple:
        return (self.stage.name, self.save, self.req, self.extra, self.drop_p)

    def drop(self, base: int) -> Optional[Bk.Drop]:
        return Bk.Drop(self.drop_p, self.seed, base) if self.drop_p > 0.0 else None


# ------------------------------------------------------------------------------------------------ CUDA graphs
class _GraphEntry:
    __slots__ = ("graph", "static_in", "outputs", "launches", "aux")


class GraphCache:
    """Captures the launch sequence of a stage body into a CUDA graph the first time a (stage, shapes, flags) key is
    seen and replays it afterwards, so a training step costs a handful of graph launches instead of ~1400 kernel
    launches issued from Python.  Inputs are copied into static buffers; outputs (including everything saved for
    the backward graph) live in the graph's private memory pool and are overwritten by the next replay."""

    timing = None  # list collecting (phase, stage signature, launches, start event, end event) when profiling

    def __init__(self, priority: int = 0):
        self.entries: Dict[tuple, _GraphEntry] = {}
        self.pool = None
        # Stream priority the graphs are captured under (kernel nodes inherit it): the trunk / transformer chain is the
        # critical path of the step, the text branch and the weight-gradient lane run next to it and should only fill
        # the SMs it leaves idle.  Measured (B200, bench step): 9.39 ms with the priority vs 9.18 ms without - the starved
        # weight-gradient lane piles up behind the chain - so it is OFF unless TOIST_GRAPH_PRIO=1.
        self.priority = priority if os.environ.get("TOIST_GRAPH_PRIO", "0") != "0" else 0
        self._stream = None

    def __deepcopy__(self, memo):
        return GraphCache(self.priority)

    def clear(self) -> None:
        self.entries.clear()
        self.pool = None

    def run(self, key: tuple, fn, inputs: Sequence[Optional[torch.Tensor]]):
        e = self.entries.get(key)
        if e is not None:
            for s_, t in zip(e.static_in, inputs):
                if s_ is not None:
                    s_.copy_(t)
            if GraphCache.timing is None:
                e.graph.replay()
            else:  # tools/stage_times.py: CUDA events around every replay (current stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                e.graph.replay()
                e1.record()
                GraphCache.timing.append((key[0], str(key[1])[:60], e.launches, e0, e1))
            K._count(e.launches)
            return e.outputs
        static_in = [None if t is None else t.detach().clone() for t in inputs]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up outside the capture: lazy kernel attributes, allocator, tile caches
            fn(*static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        n0 = K.launches()
        g = torch.cuda.CUDAGraph()
        if self.priority != 0 and (self._stream is None or self._stream.device != cur.device):
            self._stream = torch.cuda.Stream(device=cur.device, priority=self.priority)
        with torch.cuda.graph(g, pool=self.pool, stream=self._stream if self.priority != 0 else None):
            outputs = fn(*static_in)
        if self.pool is None:
            self.pool = g.pool()
        e = _GraphEntry()
        e.graph, e.static_in, e.outputs, e.launches, e.aux = g, static_in, outputs, K.launches() - n0, None
        self.entries[key] = e
        g.replay()
        return outputs


def _numel(shape) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


def _sig(tensors) -> tuple:
    return tuple(None if t is None else (tuple(t.shape), t.dtype, tuple(t.stride())) for t in tensors)


class StageFn(torch.autograd.Function):
    """Generic autograd node of one stage.  `spec.fwd(c, *inputs) -> (outputs, saved)` and
    `spec.bwd(c, saved, needs, *grad_outputs) -> (input_grads, param_grads)` are plain launch sequences (blocks.py);
    with `c.graphs` set they are captured once per shape signature and replayed."""

    @staticmethod
    def forward(ctx, spec, c: Call, *args):
        n_in = spec.n_inputs
        inputs = [a.contiguous() if isinstance(a, torch.Tensor) and spec.contiguous_inputs else a for a in args[:n_in]]
        ctx.spec, ctx.c = spec, c
        ctx.n_in = n_in
        ctx.n_tail = len(args) - n_in  # the stage's parameters, or ONE anchor tensor in direct-gradient mode
        if c.graphs is None:
            outs, saved = spec.fwd(c, *inputs)
            ctx.fkey = None
        else:
            ctx.fkey = ("fwd", c.signature(), _sig(inputs))
            outs, saved = c.graphs.run(ctx.fkey, lambda *t: spec.fwd(c, *t), inputs)
            outs = tuple(None if o is None else o.detach() for o in outs)  # fresh aliases of the static buffers
        if c.save:
            if c.graphs is None:
                _save(ctx, saved)
            else:
                # everything the backward needs is static memory of the forward graph: keep the Python structure
                # itself instead of packing ~750 tensors through save_for_backward every step
                ctx.static_saved = saved
        nd = [outs[i] for i in spec.nondiff if i < len(outs) and outs[i] is not None]
        if nd:
            ctx.mark_non_differentiable(*nd)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        spec, c = ctx.spec, ctx.c
        saved = _load(ctx) if c.graphs is None else ctx.static_saved
        needs = tuple(ctx.needs_input_grad[2: 2 + ctx.n_in])
        gouts = [None if g is None else g.contiguous() for g in gouts]
        dev = next((g.device for g in gouts if g is not None), None)
        numel = c.stage.arena_numel(c.req)

        sync = getattr(c, "grad_sync", None)
        marks_on = bool(sync is not None and c.req and sync.wants_marks())

        def body(*g):
            with Bk.zero_arena(numel, dev, marks=marks_on) as arena, K.wgrad_lanes():
                gi, gr = spec.bwd(c, saved, needs, *g)
            return gi, gr, arena.buf, tuple(arena.marks)

        params = c.stage.params if ctx.n_tail else ()
        fresh = all(p.grad is None for p in params)
        if c.graphs is None:
            gin, grads, abuf, marks = body(*gouts)
            pg = _grads_for(c.stage.names, grads, c.stage.shapes)
        else:
            key = ("bwd", ctx.fkey, needs, _sig(gouts), marks_on)
            if not fresh:
                # A gradient adopted from an earlier replay of this graph IS the graph's static buffer: the replay below
                # would overwrite what was accumulated (or zeroed in place by zero_grad(set_to_none=False)).  Detach
                # such gradients from the static memory first.
                e = c.graphs.entries.get(key)
                if e is not None:
                    static = e.outputs[1]
                    for n, p in zip(c.stage.names, params):
                        t = static.get(n)
                        if t is not None and p.grad is not None and p.grad.data_ptr() == t.data_ptr():
                            p.grad = p.grad.clone()
            gin, grads, abuf, marks = c.graphs.run(key, body, gouts)
            e = c.graphs.entries[key]
            if e.aux is None:  # the gradient buffers are static: their per-parameter views are made once per graph
                e.aux = _grads_for(c.stage.names, grads, c.stage.shapes)
            pg = e.aux
            gin = tuple(None if t is None else t.detach() for t in gin)
        if not fresh:
            # Gradient accumulation (several backward passes per optimizer step, DistributedDataParallel.no_sync()):
            # fold what the parameters already hold into this pass's arena and hand the sum over as the new gradient.
            # The all-reduce below then averages the ACCUMULATED gradient, which is what torch's wrapper exchanges on
            # the first synchronised backward after no_sync() steps.
            idx = [i for i, p in enumerate(params) if p.grad is not None and pg[i] is not None]
            if idx:
                torch._foreach_add_([pg[i] for i in idx], [params[i].grad for i in idx])
                for i in idx:
                    params[i].grad = None
        # data parallel: one in-place all-reduce of the stage's gradient arena on a side stream (util/dist.py)
        if sync is not None and c.req:
            sparse = grads.get("@sparse") if fresh else None  # accumulated gradients are dense: exchange them densely
            if sparse is not None:
                sparse = sparse + (pg[c.stage.names.index(sparse[0])],)
            # (the marks were recorded inside the stage's backward: after a fold they no longer say "final")
            sync.reduce(c.stage.name, abuf, pg, marks=marks if fresh else (), sparse=sparse)
        if ctx.n_tail == 0:
            return (None, None) + tuple(gin)
        if getattr(c, "anchor", None) is not None and ctx.n_tail == 1:
            # Direct mode: the stage owns its parameters' .grad.  Routing ~1000 parameter edges through autograd
            # (Function.apply inputs, AccumulateGrad per parameter) costs several ms of host time per step; here the
            # views of the arena are assigned in one loop and autograd only sees the anchor.
            for p, t in zip(params, pg):
                if t is not None:
                    p.grad = t
            _direct_join.note()
            return (None, None) + tuple(gin) + (None,)
        if c.graphs is not None:
            # The gradient buffers are static memory of the backward graph: hand autograd fresh aliases, which
            # AccumulateGrad adopts without a copy (every parameter's .grad is None at this point).
            pg = tuple(None if t is None else t.detach() for t in pg)
        return (None, None) + tuple(gin) + pg


class _DirectJoin:
    """Stream hygiene of direct-gradient mode.  autograd runs a node's backward on the stream its forward ran on (the
    text branch has its own) and, for gradients it accumulates itself, makes the caller's stream wait for those
    streams when backward() returns.  Gradients assigned by the stages bypass that, so every stage records an event
    on its stream and one engine callback per backward pass makes the caller's stream wait for them."""

    def __init__(self):
        self.events: List[torch.cuda.Event] = []
        self.armed = False

    def note(self) -> None:
        ev = torch.cuda.Event()
        ev.record()
        self.events.append(ev)
        if not self.armed:
            self.armed = True
            torch.autograd.Variable._execution_engine.queue_callback(self.join)

    def join(self) -> None:
        cur = torch.cuda.current_stream()
        for ev in self.events:
            cur.wait_event(ev)
        self.events.clear()
        self.armed = False


_direct_join = _DirectJoin()


class Spec:
    def __init__(self, name, n_inputs, fwd, bwd, nondiff=(), contiguous_inputs=True):
        self.name, self.n_inputs, self.fwd, self.bwd = name, n_inputs, fwd, bwd
        self.nondiff, self.contiguous_inputs = tuple(nondiff), contiguous_inputs


def run_stage(spec: Spec, c: Call, *inputs):
    if not c.req:  # no parameter of this stage wants a gradient (no_grad, eval loops, frozen detector): inputs only
        return StageFn.apply(spec, c, *inputs)
    if getattr(c, "anchor", None) is not None:
        return StageFn.apply(spec, c, *inputs, c.anchor)
    return StageFn.apply(spec, c, *inputs, *c.stage.params)




# ------------------------------------------------------------------------------------------------ backbone
_TRUNK_CHAINS = int(os.environ.get("TOIST_TRUNK_CHAINS", "2"))
# The backward data-gradient chain can be split the same way (TOIST_TRUNK_BWD_CHAINS=1); measured without gain (8.98 vs
# 8.96 ms per bench step): the weight-gradient lane already fills the chain's gaps there.  Off by default.
_TRUNK_BWD_CHAINS = os.environ.get("TOIST_TRUNK_BWD_CHAINS", "0") != "0"
_trunk_lane: Dict[torch.device, torch.cuda.Stream] = {}


def _backbone_fwd_chains(c: Call, images: torch.Tensor, chains: int):
    """The trunk forward as `chains` independent half-batch launch chains on separate streams (FrozenBatchNorm makes
    the images independent, models/backbone.py:48-58).  The trunk is a strictly serial chain of ~106 launches whose
    100..400-CTA grids leave SMs idle and whose launch / prologue / epilogue latencies are all exposed; two chains fill
    each other's gaps exactly as the weight-gradient lane does in the backward pass (measured there: 1.07 ms of 3.2).
    Every chain writes batch slices of full-batch activation tensors, so what the backward sees is unchanged."""
    st = c.stage
    w = WView(c.w, st.prefix)
    n, _, hh, ww = images.shape
    dev = images.device
    ho, wo = K.conv_out_size(hh, 7, 2, 3), K.conv_out_size(ww, 7, 2, 3)
    hp, wp = K.conv_out_size(ho, 3, 2, 1), K.conv_out_size(wo, 3, 2, 1)

    def buf(h, w_, ch):
        return torch.empty((n, h, w_, ch), dtype=BF, device=dev)

    pooled = buf(hp, wp, 64)
    plan = []  # per block: (li, bi, stride, has_ds, y1, y2, out)
    h, wd = hp, wp
    for li, (planes, nblocks) in enumerate(zip((64, 128, 256, 512), st.blocks), start=1):
        for bi in range(nblocks):
            stride = 2 if (li > 1 and bi == 0) else 1
            h2, w2 = K.conv_out_size(h, 3, stride, 1), K.conv_out_size(wd, 3, stride, 1)
            plan.append((li, bi, stride, bi == 0, buf(h, wd, planes), buf(h2, w2, planes), buf(h2, w2, planes * 4)))
            h, wd = h2, w2
    main = torch.cuda.current_stream()
    lane = _trunk_lane.get(dev)
    if lane is None:
        lane = _trunk_lane[dev] = torch.cuda.Stream(device=dev)
    lane.wait_stream(main)
    w7 = c.w.get(st.prefix + "conv1.weight7")
    per = n // chains
    with K.gemm_chains(chains):
        for ch in range(chains):
            sl = slice(ch * per, (ch + 1) * per)
            with torch.cuda.stream(main if ch == 0 else lane):
                if w7 is not None:
                    y = K.stem_conv7x7(images[sl], w7, w["bn1.shift"])
                else:
                    patches = K.stem_im2col(images[sl], STEM_LDK)
                    y = K.linear_fwd(patches, w["conv1.weight"], w["bn1.shift"], act=ACT_RELU).view(per, ho, wo, 64)
                    del patches
                x = K.maxpool3x3s2(y, out=pooled[sl])
                del y
                for (li, bi, stride, has_ds, y1, y2, out) in plan:
                    x, _ = Bk.bottleneck_fwd(w.sub(f"layer{li}.{bi}."), x, stride, has_ds, into=(y1[sl], y2[sl], out[sl]))
    main.wait_stream(lane)
    feats, saved = [], {}
    x = pooled
    for (li, bi, stride, has_ds, y1, y2, out) in plan:
        if c.save and li >= st.first_trainable:
            saved[(li, bi)] = (x, y1, y2)
        x = out
        if bi == st.blocks[li - 1] - 1:
            feats.append(out)
    return feats, saved


def backbone_fwd(c: Call, images: torch.Tensor):
    n = images.shape[0]
    if _TRUNK_CHAINS > 1 and n >= 2 * _TRUNK_CHAINS and n % _TRUNK_CHAINS == 0:
        return _backbone_fwd_chains(c, images, _TRUNK_CHAINS)
    st = c.stage
    w = WView(c.w, st.prefix)
    n, _, hh, ww = images.shape
    ho, wo = K.conv_out_size(hh, 7, 2, 3), K.conv_out_size(ww, 7, 2, 3)
    w7 = c.w.get(st.prefix + "conv1.weight7")
    if w7 is not None:
        y = K.stem_conv7x7(images, w7, w["bn1.shift"])
    else:
        patches = K.stem_im2col(images, STEM_LDK)
        y = K.linear_fwd(patches, w["conv1.weight"], w["bn1.shift"], act=ACT_RELU).view(n, ho, wo, 64)
        del patches
    x = K.maxpool3x3s2(y)
    del y
    feats, saved = [], {}
    for li, nblocks in enumerate(st.blocks, start=1):
        for bi in range(nblocks):
            stride = 2 if (li > 1 and bi == 0) else 1
            x, sv = Bk.bottleneck_fwd(w.sub(f"layer{li}.{bi}."), x, stride, bi == 0)
            if c.save and li >= st.first_trainable:
                saved[(li, bi)] = sv
        feats.append(x)
    return feats, saved


def backbone_bwd(c: Call, gfeats: Dict[int, torch.Tensor], feats: Sequence[torch.Tensor], saved) -> Dict[str, torch.Tensor]:
    """gfeats: layer index (1..4) -> gradient of that layer's output (NHWC bf16)."""
    st = c.stage
    grads: Dict[str, torch.Tensor] = {}
    w = WView(c.w, st.prefix)
    gz = None
    n = feats[-1].shape[0]
    chains = _TRUNK_CHAINS if (_TRUNK_BWD_CHAINS and _TRUNK_CHAINS > 1 and n >= 2 * _TRUNK_CHAINS
                               and n % _TRUNK_CHAINS == 0) else 1
    main = torch.cuda.current_stream()
    streams = [main]
    if chains > 1:  # the data-gradient chain as independent half-batch chains (see _backbone_fwd_chains)
        dev = feats[-1].device
        lane = _trunk_lane.get(dev)
        if lane is None:
            lane = _trunk_lane[dev] = torch.cuda.Stream(device=dev)
        streams = [main] + [lane] * (chains - 1) if chains == 2 else [main]
        chains = len(streams)
    forked = False
    for li in range(4, st.first_trainable - 1, -1):
        nblocks = st.blocks[li - 1]
        ext = gfeats.get(li)
        if ext is not None:
            if forked:  # full-batch work on the main stream: the chains meet here
                main.wait_stream(streams[1])
            ext = K.relu_bwd(ext.contiguous(), feats[li - 1])
            gz = ext if gz is None else K.add_bf16(gz, ext)
            if forked:
                streams[1].wait_stream(main)
        if gz is None:
            continue
        for bi in range(nblocks - 1, -1, -1):
            stride = 2 if (li > 1 and bi == 0) else 1
            pre = f"layer{li}.{bi}."
            need_dx = not (li == st.first_trainable and bi == 0)
            if chains > 1:
                if not forked:
                    streams[1].wait_stream(main)
                    K._WgradLane.also_wait = [streams[1]]
                    forked = True
                gz = Bk.bottleneck_bwd_chains(w.sub(pre), GView(grads, st.prefix + pre), RView(c.req, st.prefix + pre), gz,
                                              saved[(li, bi)], stride, bi == 0, need_dx, streams)
            else:
                gz = Bk.bottleneck_bwd(w.sub(pre), GView(grads, st.prefix + pre), RView(c.req, st.prefix + pre), gz,
                                       saved[(li, bi)], stride, bi == 0, need_dx)
            Bk.mark_grads()
    if forked:
        main.wait_stream(streams[1])
        if K._WgradLane.active and K._WgradLane.stream is not None:
            K._WgradLane.stream.wait_stream(streams[1])  # (the lane's last launches may still read the chain's outputs)
        K._WgradLane.also_wait = []
    return grads


def _bb_fwd(c: Call, images):
    feats, saved = backbone_fwd(c, images)
    keys = sorted(saved)
    outs = tuple(feats) if c.stage.return_interm else (feats[-1],)
    return outs, (tuple(feats), tuple(keys), tuple(saved[k] for k in keys))


def _bb_bwd(c: Call, saved, needs, *gouts):
    feats, keys, saved_list = saved
    sv = dict(zip(keys, saved_list))
    if c.stage.return_interm:
        gfeats = {i + 1: g for i, g in enumerate(gouts) if g is not None}
    else:
        gfeats = {4: gouts[0]} if gouts[0] is not None else {}
    grads = backbone_bwd(c, gfeats, feats, sv) if c.req else {}
    return (None,), grads


BACKBONE = Spec("backbone", 1, _bb_fwd, _bb_bwd)


# ------------------------------------------------------------------------------------------------ text encoder
def text_fwd(c: Call, input_ids: torch.Tensor, text_mask_u8: torch.Tensor):
    """RobertaModel.forward + FeatureResizer (eval semantics: no dropout).  Returns fp32 [L, B, d_model]."""
    st = c.stage
    w = WView(c.w, st.prefix)
    B, L = input_ids.shape
    e = w.sub("embeddings.")
    x32, pos_ids = K.embed_gather(input_ids, e["word_embeddings.weight"], e["position_embeddings.weight"],
                                  e["token_type_embeddings.weight"], st.pad_id, seq_first=True)
    x, _, m0, r0 = Bk.ln_fwd(e, "LayerNorm.", x32, st.eps)
    d_emb = c.drop(10)
    if d_emb is not None:  # RobertaEmbeddings.dropout
        K.dropout(x, d_emb.site(0), out=x)
    saved_layers = []
    for i in range(st.num_layers):
        x, sv = Bk.roberta_layer_fwd(w.sub(f"encoder.layer.{i}."), x, text_mask_u8, st.num_heads, B, st.eps,
                                     c.drop(100 + 8 * i))
        saved_layers.append(sv if c.save else None)
    r = WView(c.w, st.resizer_prefix)
    y = K.linear_fwd(x, r["fc.weight"], r["fc.bias"], out_dtype=torch.float32)
    _, out, m1, r1 = Bk.ln_fwd(r, "layer_norm.", y, 1e-12, want_bf16=False, want_f32=True)
    if c.seed is not None and st.resizer_p > 0:  # FeatureResizer.dropout (models/transformer.py:491)
        K.dropout(out, (st.resizer_p, c.seed, 20), out=out)
    saved = (input_ids, pos_ids, x32, m0, r0, tuple(saved_layers), x, y, m1, r1) if c.save else None
    return (out.view(L, B, -1),), saved


def text_bwd(c: Call, saved, needs, dout: torch.Tensor):
    st = c.stage
    grads: Dict[str, torch.Tensor] = {}
    w = WView(c.w, st.prefix)
    input_ids, pos_ids, x32, m0, r0, saved_layers, x_last, y, m1, r1 = saved
    B, L = input_ids.shape
    rp = st.resizer_prefix
    r = WView(c.w, rp)
    d = dout.contiguous().view(L * B, -1)
    if c.seed is not None and st.resizer_p > 0:
        d = K.dropout(d, (st.resizer_p, c.seed, 20))
    dy = Bk.ln_bwd(r, GView(grads, rp), RView(c.req, rp), "layer_norm.", d, y, m1, r1)
    Bk.lin_param_grads(GView(grads, rp), RView(c.req, rp), "fc.weight", "fc.bias", dy, x_last, r["fc.weight"].shape)
    body_req = any(n.startswith(st.prefix) for n in c.req)
    if not body_req:
        return (None, None), grads
    dx = K.linear_dgrad(dy, r["fc.weight"])
    for i in range(st.num_layers - 1, -1, -1):
        pre = st.prefix + f"encoder.layer.{i}."
        dx = Bk.roberta_layer_bwd(WView(c.w, pre), GView(grads, pre), RView(c.req, pre), dx, saved_layers[i],
                                  st.num_heads, B, drop=c.drop(100 + 8 * i))
        Bk.mark_grads()
    d_emb = c.drop(10)
    if d_emb is not None:
        dx = K.dropout(dx, d_emb.site(0))
    pre = st.prefix + "embeddings."
    e = WView(c.w, pre)
    dx32 = Bk.ln_bwd(e, GView(grads, pre), RView(c.req, pre), "LayerNorm.", dx, x32, m0, r0, dx_dtype=torch.float32)
    # the word table is taken LAST: it is the tail of the stage's gradient arena, which lets the data-parallel exchange
    # leave it out of the dense all-reduce and ship it as (id, row) pairs instead (util/dist.py)
    names = ("position_embeddings.weight", "token_type_embeddings.weight", "word_embeddings.weight")
    bufs = {}
    for nm in names:
        if (pre + nm) in c.req:
            t = Bk._zeros(tuple(e[nm].shape), e[nm].device)
            grads[pre + nm] = t
            bufs[nm] = t
    if bufs:
        K.embed_scatter(dx32, input_ids, pos_ids, bufs.get(names[2]), bufs.get(names[0]), bufs.get(names[1]),
                        seq_first=True, pad_id=st.pad_id)
    if names[2] in bufs:
        # rows of dx32 are l * B + b: the matching ids in the same order
        grads["@sparse"] = (pre + names[2], input_ids.t().contiguous().view(-1), dx32, st.pad_id)
    return (None, None), grads


TEXT = Spec("text", 2, text_fwd, text_bwd)


# ------------------------------------------------------------------------------------------------ encoder
def encoder_fwd(c: Call, feat: torch.Tensor, text: torch.Tensor, pos16: torch.Tensor, key_mask: torch.Tensor):
    """feat NHWC bf16 [B,h,w,C]; text fp32 [L,B,E]; pos16 bf16 [S*B,E]; key_mask uint8 [B,S].
    Outputs: img_memory fp32 [S,B,E] and the projected image rows src_proj bf16 [h*w, B, E] (mask head input)."""
    st = c.stage
    B, h, wd, _ = feat.shape
    L, _, E = text.shape
    hw = h * wd
    S = hw + L
    src = torch.empty((S * B, E), dtype=BF, device=feat.device)
    ip = WView(c.w, st.input_proj_prefix)
    Bk.seq_from_nhwc_fwd(feat, ip["weight"], ip["bias"], src[: hw * B], B)
    K.cast_bf16(text.contiguous().view(L * B, E), out=src[hw * B:])
    x = src
    xp = None  # x + pos of the next layer comes out of the previous layer's last LayerNorm launch
    saved_layers = []
    for i in range(st.num_layers):
        x, sv, xp = Bk.encoder_layer_fwd(WView(c.w, st.prefix + f"layers.{i}."), x, pos16, key_mask, st.nhead, B,
                                         c.drop(1000 + 8 * i), xp=xp, want_next_xp=i + 1 < st.num_layers)
        saved_layers.append(sv if c.save else None)
    mem = K.cast_f32(x).view(S, B, E)
    saved = (feat, (L, B, E), tuple(saved_layers)) if c.save else None
    return (mem, src[: hw * B].view(hw, B, E)), saved


def encoder_bwd(c: Call, saved, needs, dmem, dsrc_proj=None):
    st = c.stage
    feat, (L, B, E), saved_layers = saved
    _, h, wd, _ = feat.shape
    hw = h * wd
    grads: Dict[str, torch.Tensor] = {}
    d = K.cast_bf16(dmem.contiguous().view(-1, E))
    d2 = None  # gradient that reached the layer's output through the next layer's (x + pos) input
    for i in range(st.num_layers - 1, -1, -1):
        pre = st.prefix + f"layers.{i}."
        d, d2 = Bk.encoder_layer_bwd(WView(c.w, pre), GView(grads, pre), RView(c.req, pre), d, saved_layers[i], st.nhead,
                                     B, c.drop(1000 + 8 * i), dy2=d2)
    d = K.add_bf16(d, d2)  # layer 0: x and x + pos are the same tensor plus a constant
    d_img = d[: hw * B]
    if dsrc_proj is not None:  # the mask head reads src_proj (models/segmentation.py:77-78)
        d_img = K.add_bf16(d_img, dsrc_proj.contiguous().view(-1, E))
    dtext = K.cast_f32(d[hw * B:]).view(L, B, E) if needs[1] else None
    ipp = st.input_proj_prefix
    dfeat = Bk.seq_from_nhwc_bwd(GView(grads, ipp), RView(c.req, ipp), "weight", "bias", d_img, feat,
                                 c.w[ipp + "weight"], needs[0])
    return (dfeat, dtext, None, None), grads


ENCODER = Spec("encoder", 4, encoder_fwd, encoder_bwd)


# ------------------------------------------------------------------------------------------------ decoder
def decoder_fwd(c: Call, mem32: torch.Tensor, qpos32: torch.Tensor, pos16: torch.Tensor, key_mask: torch.Tensor):
    """mem32 fp32 [S,B,E], qpos32 fp32 [Q,B,E] -> hs bf16 [layers, Q*B, E] (final LayerNorm applied per layer)."""
    st = c.stage
    S, B, E = mem32.shape
    Q = qpos32.shape[0]
    mem = K.cast_bf16(mem32.contiguous().view(S * B, E))
    mem_pos = K.add_bf16(mem, pos16)
    qpos = K.cast_bf16(qpos32.contiguous().view(Q * B, E))
    tgt = torch.zeros((Q * B, E), dtype=BF, device=mem.device)
    hs = torch.empty((st.num_layers, Q * B, E), dtype=BF, device=mem.device)
    nw = WView(c.w, st.prefix + "norm.")
    saved_layers = []
    tq = None  # tgt + query_pos of the next layer comes out of the previous layer's last LayerNorm launch
    k_all = v_all = None
    wk = c.w.get(st.prefix + "cross_kv.k_weight")
    if wk is not None and wk.shape[0] == st.num_layers * E:  # models/transformer.py:394: k = memory + pos, v = memory
        k_all = K.linear_fwd(mem_pos, wk, c.w[st.prefix + "cross_kv.k_bias"])
        v_all = K.linear_fwd(mem, c.w[st.prefix + "cross_kv.v_weight"], c.w[st.prefix + "cross_kv.v_bias"])
    for i in range(st.num_layers):
        kv = None if k_all is None else (k_all[:, i * E:(i + 1) * E], v_all[:, i * E:(i + 1) * E])
        tgt, sv, tq = Bk.decoder_layer_fwd(WView(c.w, st.prefix + f"layers.{i}."), tgt, qpos, mem, mem_pos, key_mask,
                                           st.nhead, B, c.drop(2000 + 8 * i), tq=tq, want_next_tq=i + 1 < st.num_layers,
                                           kv=kv)
        _, _, m, r = K.layernorm_fwd(tgt, nw["weight"], nw["bias"], 1e-5, out16=hs[i])
        saved_layers.append((sv, tgt, m, r) if c.save else None)
    saved = ((S, B, E, Q, int(k_all is not None)), tuple(saved_layers)) if c.save else None
    return (hs,), saved


def decoder_bwd(c: Call, saved, needs, dhs):
    st = c.stage
    (S, B, E, Q, hoisted), saved_layers = saved
    grads: Dict[str, torch.Tensor] = {}
    np_ = st.prefix + "norm."
    d_next = None
    d_qpos = d_mem = None
    dk_all = dv_all = None
    if hoisted:  # gradients of the hoisted key / value projections of all layers, one column block per layer
        dev = dhs.device
        dk_all = torch.empty((S * B, st.num_layers * E), dtype=BF, device=dev)
        dv_all = torch.empty((S * B, st.num_layers * E), dtype=BF, device=dev)
    for i in range(st.num_layers - 1, -1, -1):
        sv, t3, m, r = saved_layers[i]
        dy = Bk.ln_bwd(WView(c.w, np_), GView(grads, np_), RView(c.req, np_), "", dhs[i], t3, m, r)
        pre = st.prefix + f"layers.{i}."
        dkv = (dk_all[:, i * E:(i + 1) * E], dv_all[:, i * E:(i + 1) * E]) if hoisted else None
        d_tgt, dq, dmp, dm = Bk.decoder_layer_bwd(WView(c.w, pre), GView(grads, pre), RView(c.req, pre), dy, d_next,
                                                  sv, st.nhead, B, need_tgt=i > 0, drop=c.drop(2000 + 8 * i), dkv=dkv)
        d_next = d_tgt
        d_qpos = dq if d_qpos is None else K.add_bf16(d_qpos, dq)
        if not hoisted:
            d_mem = K.add_bf16(dmp, dm) if d_mem is None else K.add_bf16(d_mem, dmp, dm)
    if hoisted and needs[0]:  # d memory = dK_all Wk_all (through memory + pos) + dV_all Wv_all: two K = layers * E GEMMs
        d_k = K.linear_dgrad(dk_all, c.w[st.prefix + "cross_kv.k_weight"])
        d_mem = K.linear_dgrad(dv_all, c.w[st.prefix + "cross_kv.v_weight"], res=d_k)
    dmem32 = K.cast_f32(d_mem).view(S, B, E) if needs[0] else None
    dqpos32 = K.cast_f32(d_qpos).view(Q, B, E) if needs[1] else None
    return (dmem32, dqpos32, None, None), grads


DECODER = Spec("decoder", 4, decoder_fwd, decoder_bwd)


# ------------------------------------------------------------------------------------------------ heads
def heads_fwd(c: Call, hs: torch.Tensor, text_mem32: Optional[torch.Tensor]):
    """hs bf16 [L, Q*B, E] -> logits [L,B,Q,C], boxes [L,B,Q,4] (sigmoid), proj_queries [L,B,Q,D], proj_tokens [B,T,D]
    (models/mdetr.py:420-433).  All four outputs are differentiable: loss_contrastive_align (models/mdetr.py:601-666,
    weight 1 per decoder layer at :1068-1069) back-propagates through both L2-normalised projections into the decoder
    states and into `text_memory` (the last L rows of the encoder output)."""
    st = c.stage
    B = c.B
    w = WView(c.w, st.prefix)
    L, QB, E = hs.shape
    Q = QB // B
    logits = Bk.heads_linear_fwd(hs, w["class_embed.weight"], w["class_embed.bias"], L, Q, B)
    hs2 = hs.view(L * QB, E)
    h1 = K.linear_fwd(hs2, w["bbox_embed.layers.0.weight"], w["bbox_embed.layers.0.bias"], act=ACT_RELU)
    h2 = K.linear_fwd(h1, w["bbox_embed.layers.1.weight"], w["bbox_embed.layers.1.bias"], act=ACT_RELU)
    boxes = Bk.heads_linear_fwd(h2.view(L, QB, E), w["bbox_embed.layers.2.weight"], w["bbox_embed.layers.2.bias"], L, Q,
                                B, act=ACT_SIGMOID)
    pq = pt = nq = nt = tm = None
    if st.contrastive:
        raw = Bk.heads_linear_fwd(hs, w["contrastive_align_projection_image.weight"],
                                  w["contrastive_align_projection_image.bias"], L, Q, B)
        D = raw.shape[-1]
        pq, nq = K.l2norm_fwd(raw.view(-1, D))  # F.normalize(p=2, dim=-1), eps 1e-12
        pq = pq.view(L, B, Q, D)
        T = text_mem32.shape[0]
        tm = K.cast_bf16(text_mem32.contiguous().view(1, T * B, E))
        rawt = Bk.heads_linear_fwd(tm, w["contrastive_align_projection_text.weight"],
                                   w["contrastive_align_projection_text.bias"], 1, T, B)
        pt, nt = K.l2norm_fwd(rawt.view(-1, D))
        pt = pt.view(B, T, D)
    saved = (hs, h1, h2, boxes, pq, nq, pt, nt, tm) if c.save else None
    return (logits, boxes, pq, pt), saved


def heads_bwd(c: Call, saved, needs, dlogits, dboxes, dpq=None, dpt=None):
    st = c.stage
    hs, h1, h2, boxes, pq, nq, pt, nt, tm = saved
    L, QB, E = hs.shape
    B = c.B
    Q = QB // B
    grads: Dict[str, torch.Tensor] = {}
    g, rq, w = GView(grads, st.prefix), RView(c.req, st.prefix), WView(c.w, st.prefix)
    dhs = None
    dtext = None
    if dboxes is not None:
        dpre = K.sigmoid_bwd(dboxes.contiguous(), boxes)
        d16 = K.cast_pad_bf16(dpre, 8)
        dh2 = Bk.heads_linear_bwd(g, rq, "bbox_embed.layers.2.weight", "bbox_embed.layers.2.bias", d16,
                                  h2.view(L, QB, E), w["bbox_embed.layers.2.weight"], L, Q, B,
                                  mask=h2.view(L, QB, E)).view(L * QB, E)
        Bk.lin_param_grads(g, rq, "bbox_embed.layers.1.weight", "bbox_embed.layers.1.bias", dh2, h1, (E, E))
        dh1 = K.linear_dgrad(dh2, w["bbox_embed.layers.1.weight"], mask=h1)
        Bk.lin_param_grads(g, rq, "bbox_embed.layers.0.weight", "bbox_embed.layers.0.bias", dh1, hs.view(L * QB, E),
                           (E, E))
        dhs = K.linear_dgrad(dh1, w["bbox_embed.layers.0.weight"]).view(L, QB, E)
    if dpq is not None and pq is not None:  # F.normalize backward, then the image projection (models/mdetr.py:432)
        D = pq.shape[-1]
        draw = K.l2norm_bwd(dpq.contiguous().view(-1, D), pq.view(-1, D), nq)
        d16 = K.cast_pad_bf16(draw.view(L, B, Q, D), (D + 7) // 8 * 8)
        dhs = Bk.heads_linear_bwd(g, rq, "contrastive_align_projection_image.weight",
                                  "contrastive_align_projection_image.bias", d16, hs,
                                  w["contrastive_align_projection_image.weight"], L, Q, B, res=dhs)
    if dpt is not None and pt is not None:  # text projection of memory_cache["text_memory"] (models/mdetr.py:433-435)
        D = pt.shape[-1]
        T = pt.shape[1]
        draw = K.l2norm_bwd(dpt.contiguous().view(-1, D), pt.view(-1, D), nt)
        d16 = K.cast_pad_bf16(draw.view(1, B, T, D), (D + 7) // 8 * 8)
        dtm = Bk.heads_linear_bwd(g, rq, "contrastive_align_projection_text.weight",
                                  "contrastive_align_projection_text.bias", d16, tm,
                                  w["contrastive_align_projection_text.weight"], 1, T, B)
        if needs[1]:
            dtext = K.cast_f32(dtm.view(T * B, E)).view(T, B, E)
    if dlogits is not None:
        d16 = K.cast_bf16(dlogits.contiguous())
        dhs = Bk.heads_linear_bwd(g, rq, "class_embed.weight", "class_embed.bias", d16, hs, w["class_embed.weight"], L,
                                  Q, B, res=dhs)
    return (dhs, dtext), grads


HEADS = Spec("heads", 2, heads_fwd, heads_bwd)
