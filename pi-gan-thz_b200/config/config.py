"""Drop-in ``config.config`` — same constant names and values as the reference's config/config.py:13-100.
Paths are rooted at the current working directory's project root (override with PIGAN_PROJECT_ROOT), not
inside the package, so a read-only install still trains."""
import os

try:
    import torch
    TORCH_AVAILABLE = True
except ImportError:  # pragma: no cover
    TORCH_AVAILABLE = False

PROJECT_ROOT = os.environ.get("PIGAN_PROJECT_ROOT", os.getcwd())

# general (config.py:16-19)
RANDOM_SEED = 42
DEVICE = "cuda" if TORCH_AVAILABLE and torch.cuda.is_available() else "cpu"
NUM_WORKERS = 4

# paths (config.py:23-33)
DATA_DIR = os.path.join(PROJECT_ROOT, "dataset")
DATASET_PATH = os.path.join(DATA_DIR, "THz_Metamaterial_Spectra_With_Metrics.csv")
FULL_DATA_PATH = DATASET_PATH
CHECKPOINT_DIR = os.path.join(PROJECT_ROOT, "checkpoints")
SAVED_MODELS_DIR = os.path.join(PROJECT_ROOT, "saved_models")
LOG_DIR = os.path.join(PROJECT_ROOT, "logs")
PLOTS_DIR = os.path.join(PROJECT_ROOT, "plots")

# data / model dimensions (config.py:37-55)
SPECTRUM_DIM = 250
NUM_SPECTRUM_POINTS = SPECTRUM_DIM
Z_DIM = 100
GENERATOR_INPUT_DIM = SPECTRUM_DIM
GENERATOR_OUTPUT_DIM = 4
GENERATOR_OUTPUT_PARAM_DIM = 4
DISCRIMINATOR_INPUT_SPEC_DIM = SPECTRUM_DIM
DISCRIMINATOR_INPUT_PARAM_DIM = 4
FORWARD_MODEL_INPUT_DIM = 4
FORWARD_MODEL_OUTPUT_SPEC_DIM = SPECTRUM_DIM
FORWARD_MODEL_OUTPUT_METRICS_DIM = 8

# training (config.py:59-74)
FWD_PRETRAIN_EPOCHS = 500
FWD_PRETRAIN_LR = 0.001
LR_FWD_SIM = 0.001
NUM_EPOCHS = 500
BATCH_SIZE = 64
LR_G = 0.0002
LR_D = 0.0002
LOG_INTERVAL = 10
SAVE_MODEL_INTERVAL = 50
SAVE_INTERVAL = 50

# generator loss weights (config.py:79-88)
LAMBDA_RECON = 100.0
LAMBDA_PHYSICS = 10.0
LAMBDA_MAXWELL = 1.0
LAMBDA_LC = 1.0
LAMBDA_PARAM_RANGE = 0.1
LAMBDA_BNN_KL = 0.0
LAMBDA_PHYSICS_SPECTRUM = 10.0
LAMBDA_PHYSICS_METRICS = 1.0


def create_directories():
    for d in (DATA_DIR, CHECKPOINT_DIR, SAVED_MODELS_DIR, LOG_DIR, PLOTS_DIR):
        os.makedirs(d, exist_ok=True)
