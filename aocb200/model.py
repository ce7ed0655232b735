"""Drop-in replacement for the reference's model module (networks/aoc/aocnet.py).

The reference eval loop builds its model with
    CFBI = importlib.import_module(cfg.MODEL_MODULE); model = CFBI.get_module()(cfg, feature_extracter).cuda(gpu)
(networks/engine/eval_manager_mm.py:39-42) and then only calls `.eval()`, `load_network` (state_dict names,
utils/checkpoint.py:49-70) and `model.forward_for_eval(...)` once per frame (eval_manager_mm.py:246-249).
Setting `cfg.MODEL_MODULE = 'aocb200.model'` selects this class instead; same constructor, same state_dict
names, same forward_for_eval contract -- every numerical op runs in hand-written sm_100a CUDA (libaocb200.so).
There is no CPU or torch fallback: on a machine without the built library or without a B200 it raises.
"""
import torch

from .engine import Engine
from .params import ParamTree


class AOCNetB200(ParamTree):
    """Same API as AOCNet (aocnet.py:11-107).  `feature_extracter` is accepted for signature parity: when it is an
    nn.Module (the reference's DeepLab) its weights are adopted under the reference's `feature_extracter.` prefix."""

    def __init__(self, cfg=None, feature_extracter=None):
        super().__init__()
        self.cfg = cfg
        self._engine = None
        self._fx = feature_extracter      # kept for forward() (training), which delegates to the reference module
        self._ref = None
        if isinstance(feature_extracter, torch.nn.Module):
            sd = {"feature_extracter." + k: v for k, v in feature_extracter.state_dict().items()}
            own = self.state_dict()
            self.load_state_dict({k: v for k, v in sd.items() if k in own and own[k].shape == v.shape}, strict=False)
        if cfg is not None:
            assert getattr(cfg, "MODEL_SEMANTIC_EMBEDDING_DIM", 100) == 100, "kernels are specialised for C=100"
            assert not getattr(cfg, "MODEL_FLOAT16_MATCHING", False)

    # -- engine lifetime: rebuilt when parameters are replaced (load_state_dict / .cuda() / .to()) -------------
    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self):
        if self._engine is None:
            dev = next(iter(self.parameters())).device
            if dev.type != "cuda":
                raise RuntimeError("AOCNetB200 runs only on a CUDA (sm_100a) device; call .cuda() first -- "
                                   "there is no CPU fallback")
            self._engine = Engine(self.state_dict(), dev)
        return self._engine

    def reserve_memory(self, nbytes, small_mb=64):
        """Serving knob (no counterpart in the reference): make torch's caching allocator hold `nbytes` of device memory
        as ONE free cached segment, from which the per-frame result tensors and the growing bank (one 10 MB embedding
        per stored frame at 480p) are then split without a cudaMalloc -- a cudaMalloc next to running graphs was measured
        at 20-100 ms, i.e. several frames, whenever the allocator had to grow in the middle of a sequence."""
        dev = next(iter(self.parameters())).device
        torch.empty(int(nbytes), dtype=torch.uint8, device=dev)      # freed at once: stays cached, splittable
        # tensors of <= 1 MB (the uint8 label maps a sequence keeps per stored frame) come from the allocator's SMALL pool,
        # 2 MB segments of their own: the one cudaMalloc left inside a 26-frame run was such a segment (allocator trace)
        small = [torch.empty(1 << 20, dtype=torch.uint8, device=dev) for _ in range(int(small_mb))]
        del small

    # -- reference API ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_for_eval(self, memory_prev_list, ref_embeddings, ref_masks, prev_embedding, prev_mask, current_frame,
                         pred_size, gt_ids):
        """aocnet.py:84-107.  Returns (probs [1,K+1,H,W] | None, embedding [1,100,h,w], memory list)."""
        return self.engine().forward_for_eval(memory_prev_list, ref_embeddings, ref_masks, prev_embedding, prev_mask,
                                              current_frame, pred_size, gt_ids)

    @torch.no_grad()
    def extract_feature(self, x):
        """aocnet.py:109-112 -> (embedding [1,100,h,w], low_level [1,256,h,w])"""
        emb, low = self.engine().extract_feature(x)
        return emb.nchw(), low.nchw()

    def forward(self, input, memory_prev_list=None, ref_frame_label=None, previous_frame_mask=None,
                current_frame_mask=None, gt_ids=None, step=0, tf_board=False):
        """aocnet.py:54-82 is the TRAINING forward (loss + boards).  Training is outside this inference engine, so the
        call is delegated to the reference's own torch module: it is built on first use from the
        reference package on sys.path (`networks.aoc.aocnet`, or cfg.MODEL_REFERENCE_MODULE), around the
        `feature_extracter` this object was constructed with, and SHARES this module's parameter tensors
        (load_state_dict(assign=True)), so an optimiser stepping either side updates both.  Without the reference
        package there is nothing to delegate to and the call raises."""
        return self._reference_module()(input, memory_prev_list, ref_frame_label, previous_frame_mask,
                                        current_frame_mask, gt_ids, step=step, tf_board=tf_board)

    def _reference_module(self):
        if self._ref is None:
            import importlib
            name = getattr(self.cfg, "MODEL_REFERENCE_MODULE", "networks.aoc.aocnet")
            try:
                mod = importlib.import_module(name)
            except ImportError as e:
                raise NotImplementedError(
                    "AOCNetB200.forward (training) delegates to the reference's torch module '%s', which is not "
                    "importable here (%s); aocb200 itself accelerates forward_for_eval only" % (name, e)) from e
            if not isinstance(self._fx, torch.nn.Module):
                raise NotImplementedError("AOCNetB200.forward needs the feature_extracter module it was constructed with")
            ref = mod.get_module()(self.cfg, self._fx)
            ref.load_state_dict(self.state_dict(keep_vars=True), strict=False, assign=True)
            self._ref = ref
        return self._ref


def get_module():
    """networks/aoc/aocnet.py:374-375"""
    return AOCNetB200
