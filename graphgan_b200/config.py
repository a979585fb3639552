"""Run-time configuration: the reference's ``config`` surface, name for name and default for default.

The reference keeps its hyper-parameters as attributes of a flat module that every layer reads at CALL
time (``config.<name>``; /root/reference src/GraphGAN/config.py:1-41), so callers and tests can patch
them.  This module keeps that contract: the attributes below exist under the same names with the same
values, and B200-only knobs are appended with defaults that leave the reference behaviour unchanged.
``src/GraphGAN/config.py`` aliases this module so ``import config`` keeps working from that directory.
"""


def _install(table):
    globals().update(table)


# -- optimisation schedule (reference config.py:4-16) ---------------------------------------------------
_install(dict(
    modes=["gen", "dis"],        # order of the two models in the result / embedding files
    batch_size_gen=64, batch_size_dis=64,          # pairs per optimizer step
    lambda_gen=1e-5, lambda_dis=1e-5,              # l2 weights of the two losses
    n_sample_gen=20,                               # walks per root in a generator pass
    lr_gen=1e-3, lr_dis=1e-3,                      # Adam learning rates
    n_epochs=20, n_epochs_gen=30, n_epochs_dis=30,  # outer loop / inner loops per epoch
    update_ratio=1,                                # fraction of roots resampled per pass
))
gen_interval = n_epochs_gen      # noqa: F821  resample generator data every this many inner epochs
dis_interval = n_epochs_dis      # noqa: F821  same for the discriminator

# -- checkpointing and model shape (reference config.py:19-25) ------------------------------------------
_install(dict(load_model=False, save_steps=10, n_emb=50, multi_processing=False, window_size=2))

# -- task, dataset and the nine path strings (reference config.py:28-41), relative to src/GraphGAN ------
app, dataset = "link_prediction", "CA-GrQc"


def _paths(app_name, data_name):
    data, pre, res = "../../data/" + app_name + "/", "../../pre_train/" + app_name + "/", "../../results/" + app_name + "/"
    emb = pre + data_name + "_pre_train.emb"
    return dict(
        train_filename=data + data_name + "_train.txt",
        test_filename=data + data_name + "_test.txt",
        test_neg_filename=data + data_name + "_test_neg.txt",
        pretrain_emb_filename_d=emb, pretrain_emb_filename_g=emb,
        emb_filenames=[res + data_name + "_gen_.emb", res + data_name + "_dis_.emb"],
        result_filename=res + data_name + ".txt",
        cache_filename="../../cache/" + data_name + ".pkl",     # unused here: trees are rebuilt on the GPU
        model_log="../../log/",
    )


_install(_paths(app, dataset))

# -- B200 additions (not in the reference) ---------------------------------------------------------------
_install(dict(
    device="cuda:0",            # one process per GPU; LOCAL_RANK overrides the index under torchrun
    seed=0,                     # Philox key of the walk sampler and seed of the batch shuffles
    root_batch=4096,            # roots whose BFS parent arrays are resident at once (4*N bytes each)
    tree_cache_bytes=8 << 30,   # keep ALL trees resident (like the reference's cache) below this size
    max_path_len=64,            # row stride of recorded generator paths; a longer walk is an error
))
