"""Configuration surface of the reference (src/GraphGAN/config.py:1-41), kept name for name and
default for default: a flat module read as ``config.<name>`` at call time, so callers (and tests)
may monkey-patch it exactly as they could with the reference.

GPU-only knobs are ADDED at the bottom with defaults that leave the reference behaviour
unchanged.  Paths stay relative to the working directory ``src/GraphGAN`` like the reference's.
"""
modes = ["gen", "dis"]

# training settings (config.py:4-16)
batch_size_gen = 64  # batch size for the generator
batch_size_dis = 64  # batch size for the discriminator
lambda_gen = 1e-5  # l2 loss regulation weight for the generator
lambda_dis = 1e-5  # l2 loss regulation weight for the discriminator
n_sample_gen = 20  # number of samples for the generator
lr_gen = 1e-3  # learning rate for the generator
lr_dis = 1e-3  # learning rate for the discriminator
n_epochs = 20  # number of outer loops
n_epochs_gen = 30  # number of inner loops for the generator
n_epochs_dis = 30  # number of inner loops for the discriminator
gen_interval = n_epochs_gen  # sample new nodes for the generator for every gen_interval iterations
dis_interval = n_epochs_dis  # sample new nodes for the discriminator for every dis_interval iterations
update_ratio = 1  # updating ratio when choose the trees

# model saving (config.py:19-20)
load_model = False  # whether loading existing model for initialization
save_steps = 10

# other hyper-parameters (config.py:23-25)
n_emb = 50
multi_processing = False  # kept for source compatibility; trees are built on the GPU
window_size = 2

# application and dataset settings (config.py:28-29)
app = "link_prediction"
dataset = "CA-GrQc"

# path settings (config.py:32-41)
train_filename = "../../data/" + app + "/" + dataset + "_train.txt"
test_filename = "../../data/" + app + "/" + dataset + "_test.txt"
test_neg_filename = "../../data/" + app + "/" + dataset + "_test_neg.txt"
pretrain_emb_filename_d = "../../pre_train/" + app + "/" + dataset + "_pre_train.emb"
pretrain_emb_filename_g = "../../pre_train/" + app + "/" + dataset + "_pre_train.emb"
emb_filenames = ["../../results/" + app + "/" + dataset + "_gen_.emb",
                 "../../results/" + app + "/" + dataset + "_dis_.emb"]
result_filename = "../../results/" + app + "/" + dataset + ".txt"
cache_filename = "../../cache/" + dataset + ".pkl"
model_log = "../../log/"

# ---------------------------------------------------------------------------------------------
# B200 additions (not in the reference).  Defaults keep the reference semantics.
# ---------------------------------------------------------------------------------------------
device = "cuda:0"       # one process per GPU; under torchrun LOCAL_RANK overrides the index
seed = 0                # Philox key of the walk sampler; numpy RandomState seed of the batch shuffles
root_batch = 4096       # roots whose BFS parent arrays are resident at once (4*N bytes each)
tree_cache_bytes = 8 << 30   # keep ALL trees resident (like the reference's pickle cache) below this size
max_path_len = 64       # row stride of the recorded G paths (root .. sample, father); overflow is an error
