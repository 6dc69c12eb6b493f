"""Drop-in installer: makes the reference's OWN scripts (test_dice.py, train_onecube.py) run on the B200 path.

    import neuroclear_b200.dropin as dropin
    dropin.install()                # once, before `from models import create_model` etc. are used

With the reference's package root on sys.path, install() rebinds — inside the reference's modules — exactly the names
INTEGRATION.md lists; everything else (options, TestModel / BaseModel, the apollo model class with its set_input /
forward / backward_G / backward_D_* / optimize_parameters, the script's loops) is the reference's code, unmodified:

    models/networks.py        define_G, define_D, GANLoss, get_scheduler   -> neuroclear_b200.networks / .discriminator
    data/diceImage_dataset.py DiceImageDataSet                             -> neuroclear_b200.dicing (as a BaseDataset)
    data/singlevolume_dataset.py SingleVolumeDataset                       -> neuroclear_b200.augment (as a BaseDataset)
    util/assemble_dice.py     Assemble_Dice                                -> neuroclear_b200.dicing
    models/axial_to_lateral_gan_{apollo,dryops}_model.py  Volume           -> neuroclear_b200.projection
    test_dice.py:46,151       tifffile.imsave / skimage.io.imread of (Z,Y,X) volumes -> neuroclear_b200.volume_io

The reference finds datasets by class name among BaseDataset subclasses (data/__init__.py:20-40), so the mirrors are
re-based on the reference's BaseDataset here.  The optimisers stay torch.optim.Adam as written in the reference model
(our modules expose ordinary nn.Parameters with .grad); neuroclear_b200.apollo_model uses the fused Adam instead.
"""
from __future__ import annotations

import importlib
import sys

_INSTALLED = False


def _rebase(mirror, base_dataset, name):
    """class <name>(BaseDataset, mirror): found by data.find_dataset_using_name, constructed as cls(opt)"""
    def __init__(self, opt, *a, **kw):
        base_dataset.__init__(self, opt)
        mirror.__init__(self, opt, *a, **kw)
    ns = {"__init__": __init__, "__doc__": mirror.__doc__,
          "modify_commandline_options": staticmethod(getattr(mirror, "modify_commandline_options",
                                                             lambda parser, is_train=False: parser))}
    for meth in ("__len__", "__getitem__"):            # abstract in BaseDataset: bind the mirror's explicitly
        ns[meth] = getattr(mirror, meth)
    return type(name, (base_dataset, mirror), ns)


def install(training_data_on_gpu: bool = True):
    """Idempotent.  Requires the reference root (or oracle/_ref/neuroclear.zip) on sys.path."""
    global _INSTALLED
    if _INSTALLED:
        return
    from . import augment, dicing, discriminator, networks, projection, volume_io

    ref_networks = importlib.import_module("models.networks")
    for name in ("define_G", "get_scheduler", "init_net", "init_weights", "get_norm_layer"):
        setattr(ref_networks, name, getattr(networks, name))
    ref_networks.define_D = discriminator.define_D
    ref_networks.GANLoss = discriminator.GANLoss

    base = importlib.import_module("data.base_dataset").BaseDataset
    ref_dice = importlib.import_module("data.diceImage_dataset")
    ref_dice.DiceImageDataSet = _rebase(dicing.DiceImageDataSet, base, "DiceImageDataSet")

    if training_data_on_gpu:
        class _SingleVolume(augment.SingleVolumeDataset):
            """reads the one training volume like data/singlevolume_dataset.py:31-32, then keeps it in HBM"""
            def __init__(self, opt):
                from skimage import io
                path = importlib.import_module("data.image_folder").make_dataset(opt.dataroot, 1)[0]
                augment.SingleVolumeDataset.__init__(self, opt, io.imread(path))
                self.A_path = path
        ref_single = importlib.import_module("data.singlevolume_dataset")
        ref_single.SingleVolumeDataset = _rebase(_SingleVolume, base, "SingleVolumeDataset")

    ref_asm = importlib.import_module("util.assemble_dice")

    class Assemble_Dice(dicing.Assemble_Dice):
        """the reference builds its own dataset inside Assemble_Dice(opt) (assemble_dice.py:13-16): same here"""
        def __init__(self, opt):
            dicing.Assemble_Dice.__init__(self, opt, ref_dice.DiceImageDataSet(opt))
    ref_asm.Assemble_Dice = Assemble_Dice

    for mod in ("models.axial_to_lateral_gan_apollo_model", "models.axial_to_lateral_gan_dryops_model"):
        importlib.import_module(mod).Volume = projection.Volume

    # volume I/O of the scripts: multi-page TIFF straight from / into pinned planes
    tifffile = sys.modules.get("tifffile")
    if tifffile is not None:
        tifffile.imsave = lambda path, volume, **kw: volume_io.write_volume(path, volume)
    _INSTALLED = True
