"""Runs one of the REFERENCE's own scripts (byte-compiled in oracle/_ref/neuroclear.zip by oracle/build_ref.py) —
test_dice.py or train_onecube.py — either as shipped on the CPU (`--stock`) or on the B200 path through
neuroclear_b200.dropin.install().  Test infrastructure (used by tests/test_gpu_dropin.py).

    python tests/run_reference_script.py test_dice [--stock] -- <the script's own command line>
    python tests/run_reference_script.py train_onecube [--stock] [--iters 2] -- <command line>

Third-party modules the reference imports and this image lacks are stubbed here (they are UI / file-format helpers,
SURVEY.md §8c): skimage (io.imread -> .npy / multi-page TIFF reader; exposure.rescale_intensity -> the oracle's
restatement), tifffile.imsave, dominate, matplotlib.  train_onecube.py loops forever: the runner stops it after
--iters iterations from inside model.update_learning_rate and prints the losses of every iteration as JSON.
"""
import json
import os
import sys
import types
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Stop(Exception):
    pass


def _stub_modules():
    from neuroclear_b200 import volume_io
    from oracle import assemble as _asm, postprocess as _post
    if not hasattr(np, "float"):
        np.float = float

    def imread(path):
        return np.load(path) if str(path).endswith(".npy") else volume_io.read_volume(path)

    def imsave(path, volume, **kw):
        volume_io.write_volume(path, np.asarray(volume))

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    io_ = mod("skimage.io", imread=imread)
    ex = mod("skimage.exposure", rescale_intensity=lambda image, in_range: _asm.rescale_intensity(image, in_range),
             match_histograms=_post.match_histograms)
    tr = mod("skimage.transform")
    mod("skimage", io=io_, exposure=ex, transform=tr)
    mod("tifffile", imsave=imsave)

    class _Tag:
        def __init__(self, *a, **k):
            pass
        def __enter__(self):
            return self
        def __exit__(self, *a):
            return False
        def add(self, *a, **k):
            return self
        def __iadd__(self, other):
            return self
        def render(self):
            return ""
    tags = mod("dominate.tags", **{t: _Tag for t in ("meta", "h3", "table", "tr", "td", "p", "a", "img", "br")})
    dom = mod("dominate", tags=tags)
    dom.document = lambda *a, **k: types.SimpleNamespace(head=_Tag(), add=lambda *x, **y: None, render=lambda: "",
                                                         __enter__=None)
    class _Doc(_Tag):
        head = _Tag()
    dom.document = _Doc
    plt = mod("matplotlib.pyplot")
    mpl = mod("matplotlib", pyplot=plt, cm=mod("matplotlib.cm"))
    mod("mpl_toolkits.mplot3d", Axes3D=object)
    mod("mpl_toolkits", mplot3d=sys.modules["mpl_toolkits.mplot3d"])
    return mpl


def main():
    argv = sys.argv[1:]
    script = argv.pop(0)
    stock = "--stock" in argv[:argv.index("--")]
    iters = 2
    head = argv[:argv.index("--")]
    if "--iters" in head:
        iters = int(head[head.index("--iters") + 1])
    script_args = argv[argv.index("--") + 1:]
    archive = os.path.join(ROOT, "oracle", "_ref", "neuroclear.zip")
    if not os.path.exists(archive):
        sys.exit("oracle/_ref/neuroclear.zip is missing: python -m oracle.build_ref (needs /root/reference)")
    _stub_modules()
    sys.path.insert(0, archive)
    if not stock:
        import neuroclear_b200.dropin as dropin
        dropin.install()
    import random
    import torch
    torch.manual_seed(0)            # the scripts seed nothing: fix every stream so that two runs are comparable
    np.random.seed(0)
    random.seed(0)
    import util.html as html

    class HTML:                                        # dominate page builder: UI, not part of the path
        def __init__(self, web_dir, title, refresh=0):
            self.web_dir, self.img_dir = web_dir, os.path.join(web_dir, "images")
            os.makedirs(self.img_dir, exist_ok=True)

        def get_image_dir(self):
            return self.img_dir

        def __getattr__(self, name):
            return lambda *a, **k: None
    html.HTML = HTML
    losses = []
    if script == "train_onecube":
        import util.visualizer as vis

        class Visualizer:                              # tensorboard / HTML UI: not part of the path
            def __init__(self, opt):
                pass
            def __getattr__(self, name):
                return lambda *a, **k: None
        vis.Visualizer = Visualizer
        from models.base_model import BaseModel
        orig = BaseModel.update_learning_rate

        def counted(self):
            orig(self)
            losses.append({k: float(v) for k, v in self.get_current_losses().items()})
            if len(losses) >= iters:
                raise _Stop()
        BaseModel.update_learning_rate = counted
    code = None
    with zipfile.ZipFile(archive) as z:
        import marshal
        code = marshal.loads(z.read(script + ".pyc")[16:])
    sys.argv = [script + ".py"] + script_args
    g = {"__name__": "__main__", "__file__": script + ".py"}
    try:
        exec(code, g)
    except _Stop:
        pass
    if losses:
        print("LOSSES_JSON " + json.dumps(losses))


if __name__ == "__main__":
    main()
