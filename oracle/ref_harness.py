"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Level-2 oracle harness: imports the *unmodified* reference
(`/root/reference`, read-only) on CPU behind a table of stubs for the third-party
packages that are absent from this image, so that

    notebooks/infer.py:62   InferenceWrapper.__init__ / forward (:355-647)
    models/stage_1/volumetric_avatar/va.py:39  Model

run end to end with random-init weights.  It is used ONLY to (a) generate the golden
fixtures under tests/golden/ (see oracle/make_golden.py) and (b) pin the standalone
restatement in oracle/restatement.py.  /root/reference does not exist on the GPU box,
so nothing under tests -m gpu, smoke() or bench.py imports this file.

Stub table (what, and the reference line that needs it):
  apex                              notebooks/infer.py:12
  mediapipe, facenet_pytorch        notebooks/infer.py:21-22
  repos.MODNet.src.models.modnet    notebooks/infer.py:18, va_losses_and_visuals.py:10
  repos.face_par_off.model.BiSeNet  networks/volumetric_avatar/face_parcing.py:22
  repos.resnet                      networks/volumetric_avatar/expression_embedder.py:17
  face_alignment.detection.sfd      notebooks/infer.py:155
  skimage.measure                   models/stage_1/volumetric_avatar/va.py:12
  ibug.*                            va.py:24-27, utils/non_specific.py:8-11
  lmdb, albumentations              datasets/voxceleb2hq_pairs.py:1,6
  lpips, pytorch_msssim, wandb ...  losses/*
  EmoPortraits.networks alias       models/stage_1/volumetric_avatar/va_arguments.py:5
  datasets.voxceleb2hq_pairs        shadowed by HF `datasets`; loaded by path
Patches: torchvision resnet(pretrained=True)->weights=None; torch.load of missing files
-> {}; load_state_dict({}) tolerated; Face_vector / Face_vector_resnet (download VGG-Face
in their ctor, networks/volumetric_avatar/utils.py:1380,1471) replaced by inert objects;
every hard-coded 'cuda' redirected to CPU (`Tensor.cuda`, `Module.cuda` -> identity).
"""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import os
import pathlib
import sys
import tempfile
import types

import numpy as np
import torch
from torch import nn

REF = pathlib.Path(os.environ.get("EMO_REFERENCE_DIR", "/root/reference"))


def available() -> bool:
    return (REF / "notebooks" / "infer.py").exists()


class _Anything:
    """Inert object: any attribute/call returns another inert object."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


def _stub(name: str) -> types.ModuleType:
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = _StubModule(n)
            m.__path__ = []  # behave as a package
            m.__spec__ = importlib.machinery.ModuleSpec(n, None, is_package=True)
            sys.modules[n] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[: i - 1])], parts[i - 1], m)
    return sys.modules[name]


class _OnesMaskNet(nn.Module):
    """Stand-in for MODNet / BiSeNet: the mask networks are OUT OF SCOPE;
    returns an all-ones matte so masking is the identity."""

    def __init__(self, *a, **k):
        super().__init__()
        self.n_classes = k.get("n_classes", 19)

    def forward(self, x, *a, **k):
        b, _, h, w = x.shape
        ones = torch.ones(b, 1, h, w, dtype=x.dtype, device=x.device)
        return ones, ones, ones


class _BiSeNet(nn.Module):
    """Stand-in BiSeNet: logits whose argmax is class 1 ('face') everywhere -> mask of ones."""

    def __init__(self, n_classes=19, *a, **k):
        super().__init__()
        self.n_classes = n_classes

    def forward(self, x):
        b, _, h, w = x.shape
        y = torch.zeros(b, self.n_classes, h, w, dtype=x.dtype, device=x.device)
        y[:, 1] = 1.0
        return (y,)


_INSTALLED = False


def install_stubs():
    """Idempotently prepare sys.modules / patches so the reference imports on a CPU-only box."""
    global _INSTALLED
    if _INSTALLED:
        return
    assert available(), f"reference tree not found at {REF}"
    import torchvision  # noqa: F401  (must be fully imported before any stub lands in sys.modules)
    import torchvision.models  # noqa: F401
    import scipy.linalg  # noqa: F401
    for name in [
        "apex", "apex.parallel", "mediapipe", "mediapipe.solutions", "mediapipe.solutions.face_detection",
        "facenet_pytorch", "repos", "repos.MODNet", "repos.MODNet.src", "repos.MODNet.src.models",
        "repos.MODNet.src.models.modnet", "repos.face_par_off", "repos.face_par_off.model", "repos.resnet",
        "face_alignment", "face_alignment.detection", "face_alignment.detection.sfd",
        "ibug", "ibug.face_detection", "ibug.face_parsing", "ibug.face_parsing.utils", "ibug.roi_tanh_warping",
        "lmdb", "albumentations", "albumentations.pytorch", "lpips", "pytorch_msssim", "wandb", "kornia",
    ]:
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
    for name in ["skimage", "skimage.measure", "matplotlib", "matplotlib.pyplot", "cv2", "sklearn"]:
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)

    sys.modules["repos.MODNet.src.models.modnet"].MODNet = _OnesMaskNet
    sys.modules["repos.face_par_off.model"].BiSeNet = _BiSeNet
    sys.modules["repos.resnet"].ResNet18 = _Anything

    # --- make the reference tree importable, ahead of site-packages' `datasets`/`utils`
    sys.path.insert(0, str(REF))
    for shadow in ["datasets", "utils", "networks", "models", "losses"]:
        m = sys.modules.get(shadow)
        if m is not None and not str(getattr(m, "__file__", "") or "").startswith(str(REF)):
            del sys.modules[shadow]
            for k in [k for k in sys.modules if k.startswith(shadow + ".")]:
                del sys.modules[k]
    # `datasets` in the reference is a namespace dir without __init__; force it to resolve there
    ds = types.ModuleType("datasets")
    ds.__path__ = [str(REF / "datasets")]
    sys.modules["datasets"] = ds
    # EmoPortraits.networks alias (va_arguments.py:5)
    emo = types.ModuleType("EmoPortraits")
    emo.__path__ = [str(REF)]
    sys.modules["EmoPortraits"] = emo

    # --- CPU redirection and missing-file tolerance
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.manual_seed_all = lambda *a, **k: None
    _orig_to = nn.Module.to

    def _to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return _orig_to(self, *a, **k)

    nn.Module.to = _to
    _orig_tto = torch.Tensor.to

    def _tto(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return _orig_tto(self, *a, **k)

    torch.Tensor.to = _tto

    _orig_load = torch.load

    def _load(f, *a, **k):
        try:
            if isinstance(f, (str, os.PathLike)) and not os.path.exists(f):
                return {}
        except Exception:
            pass
        k.setdefault("weights_only", False)
        return _orig_load(f, *a, **k)

    torch.load = _load
    _orig_lsd = nn.Module.load_state_dict

    def _lsd(self, sd, *a, **k):
        if isinstance(sd, dict) and len(sd) == 0:
            return None
        return _orig_lsd(self, sd, *a, **k)

    nn.Module.load_state_dict = _lsd
    nn.DataParallel = lambda m, *a, **k: m

    import torchvision.models as tvm

    for name in ["resnet18", "resnet34", "resnet50", "vgg16", "vgg19"]:
        orig = getattr(tvm, name)

        def mk(orig):
            def f(pretrained=False, **k):
                k.pop("weights", None)
                return orig(weights=None, **k)

            return f

        setattr(tvm, name, mk(orig))

    _INSTALLED = True


def _patch_loss_only_classes():
    import networks.volumetric_avatar.utils as vu

    vu.Face_vector = _Anything
    vu.Face_vector_resnet = _Anything
    import networks.volumetric_avatar as va_pkg

    va_pkg.utils.Face_vector = _Anything
    va_pkg.utils.Face_vector_resnet = _Anything
    import utils.non_specific as ns

    ns.FaceParsingBUG = _Anything


def reference_args_lines(image_size: int = 512) -> list[str]:
    """Turn experiments/args.txt (one launch line) + every argparse default into `k: v` lines, the
    format train.py:80-83 writes and utils/args.py:54-65 re-parses."""
    install_stubs()
    import argparse

    launch = (REF / "experiments" / "args.txt").read_text().split()
    argv = launch[launch.index("../train.py") + 1:]
    parser = argparse.ArgumentParser(conflict_handler="resolve")
    parser.add = parser.add_argument
    # train.py:478-531 top-level flags that the model reads
    from utils import args as args_utils

    spec = importlib.util.spec_from_file_location("_ref_train_flags", REF / "train.py")
    src = (REF / "train.py").read_text()
    # extract only the parser.add(...) lines of train.py's __main__ block (no execution of Trainer)
    flag_lines = [l.strip() for l in src.splitlines() if l.strip().startswith("parser.add(") or
                  l.strip().startswith("parser.add_argument(")]
    env = {"parser": parser, "args_utils": args_utils, "str": str, "int": int, "float": float}
    for l in flag_lines:
        try:
            exec(l, env)
        except Exception:
            pass
    vp = importlib.util.spec_from_file_location("_ref_vox", REF / "datasets" / "voxceleb2hq_pairs.py")
    vox = importlib.util.module_from_spec(vp)
    sys.modules["datasets.voxceleb2hq_pairs"] = vox
    vp.loader.exec_module(vox)
    parser = vox.DataModule.add_argparse_args(parser)
    from models.stage_1.volumetric_avatar.va_arguments import VolumetricAvatarConfig

    parser = VolumetricAvatarConfig.add_argparse_args(parser)
    args, _ = parser.parse_known_args(argv)
    d = vars(args)
    d["image_size"] = image_size
    d["aug_warp_size"] = image_size
    d["num_gpus"] = 0
    return [f"{k}: {v}\n" for k, v in sorted(d.items())]


SANE_POSE_BIAS = [1, 1, 1, .15, -.1, .05, .03, -.02, .01]


def build_reference_wrapper(image_size: int = 512, seed: int = 0, workdir: str | None = None):
    """Instantiate the reference InferenceWrapper on CPU with seeded random-init weights.
    Returns (wrapper, model_state_dict, head_pose_state_dict, args_txt_lines)."""
    install_stubs()
    work = pathlib.Path(workdir or tempfile.mkdtemp(prefix="emo_oracle_"))
    exp = "oracle_exp"
    (work / "logs" / exp / "checkpoints").mkdir(parents=True, exist_ok=True)
    if not (work / "data").exists():
        os.symlink(REF / "data", work / "data")
    lines = reference_args_lines(image_size)
    (work / "logs" / exp / "args.txt").write_text("".join(lines))

    _patch_loss_only_classes()
    sys.argv = [sys.argv[0]]
    torch.manual_seed(seed)
    np.random.seed(seed)
    infer = importlib.import_module("notebooks.infer")
    import models.stage_1.volumetric_avatar.va as va_mod

    va_mod.FaceParsingBUG = _Anything
    w = infer.InferenceWrapper(experiment_name=exp, model_file_name="none.pth", use_gpu=False, num_gpus=0,
                               project_dir=str(work), folder="logs", print_params=False)
    # sane pose: otherwise random theta is near-singular and every sample lands out of bounds
    net = w.model.head_pose_regressor.net
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        net.fc.weight.mul_(0.01)
        net.fc.bias.copy_(torch.tensor(SANE_POSE_BIAS, dtype=torch.float32))
    net.eval()
    w.model.eval()
    msd = {k: v.detach().clone() for k, v in w.model.state_dict().items()}
    hsd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return w, msd, hsd, lines


def synthetic_frame(image_size: int, seed: int):
    """uint8 RGB noise frame -> PIL (RandomState(seed).rand(H,W,3)*255)."""
    from PIL import Image

    a = (np.random.RandomState(seed).rand(image_size, image_size, 3) * 255).astype(np.uint8)
    return Image.fromarray(a)


# ------------------------------------------------------------------------------------------------------------------
# stage 2 (notebooks/infer_s2.py:53-387, models/stage_2/base/volumetric_avatar_two.py)
# ------------------------------------------------------------------------------------------------------------------
def reference_args_lines_s2(output_size: int = 512) -> list[str]:
    """Stage-2 `args.txt` lines from the argparse defaults (the shipped stage-2 args.txt is a download, README.md:127-129):
    train.py flags + DataModule flags + models/stage_2/base/volumetric_avatar_two.py Model.add_argparse_args."""
    install_stubs()
    import argparse

    from utils import args as args_utils

    parser = argparse.ArgumentParser(conflict_handler="resolve")
    parser.add = parser.add_argument
    src = (REF / "train.py").read_text()
    env = {"parser": parser, "args_utils": args_utils, "str": str, "int": int, "float": float}
    for l in [l.strip() for l in src.splitlines() if l.strip().startswith(("parser.add(", "parser.add_argument("))]:
        try:
            exec(l, env)
        except Exception:
            pass
    if "datasets.voxceleb2hq_pairs" not in sys.modules:
        vp = importlib.util.spec_from_file_location("_ref_vox", REF / "datasets" / "voxceleb2hq_pairs.py")
        vox = importlib.util.module_from_spec(vp)
        sys.modules["datasets.voxceleb2hq_pairs"] = vox
        vp.loader.exec_module(vox)
    parser = sys.modules["datasets.voxceleb2hq_pairs"].DataModule.add_argparse_args(parser)
    _stub("datasets.Retinaface")
    two = importlib.import_module("models.stage_2.base.volumetric_avatar_two")
    parser = two.Model.add_argparse_args(parser)
    args, _ = parser.parse_known_args([])
    d = vars(args)
    d.update(output_size_s2=output_size, num_gpus=0, model_name="volumetric_avatar")
    return [f"{k}: {v}\n" for k, v in sorted(d.items())]


def build_reference_stage2(output_size: int = 512, seed: int = 0, workdir: str | None = None):
    """Reference stage-2 InferenceWrapper on CPU, seeded random init.  Returns (wrapper, state_dict, args lines)."""
    install_stubs()
    work = pathlib.Path(workdir or tempfile.mkdtemp(prefix="emo_oracle_s2_"))
    exp = "oracle_exp_s2"
    (work / "logs_s2" / exp / "checkpoints").mkdir(parents=True, exist_ok=True)
    lines = reference_args_lines_s2(output_size)
    (work / "logs_s2" / exp / "args.txt").write_text("".join(lines))
    sys.argv = [sys.argv[0]]
    torch.manual_seed(seed)
    infer_s2 = importlib.import_module("notebooks.infer_s2")
    w = infer_s2.InferenceWrapper(experiment_name=exp, model_file_name="none.pth", use_gpu=False, num_gpus=0,
                                  project_dir=str(work))
    w.model_two.eval()
    sd = {k: v.detach().clone() for k, v in w.model_two.state_dict().items()}
    return w, sd, lines
