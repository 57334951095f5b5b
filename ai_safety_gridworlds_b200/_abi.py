"""ctypes mirror of include/gwsim.h and the loader of the in-tree CUDA library.

There is NO CPU fallback: `load()` raises if libgwsim.so is missing or does not export the
ABI this module expects, and every compute entry point returns GW_ERR_NO_DEVICE without a GPU.
"""
import ctypes as C
import os

GW_ABI_VERSION = 1
GW_MAX_CELLS = 64
GW_MAX_LAYERS = 16
GW_MAX_REWARDS = 12
GW_MAX_EVENTS = 16
GW_MAX_METRICS = 16
GW_STATE_WORD_BYTES = 16
GW_STATS_LEN = 8 + GW_MAX_REWARDS
GW_STATS_RAW_LEN = 32

GW_OK, GW_ERR_INVALID, GW_ERR_CUDA, GW_ERR_NO_DEVICE, GW_ERR_STATE = 0, 1, 2, 3, 4

GW_ENV_ISLAND_NAVIGATION_EX = 1
GW_ENV_BOAT_RACE_EX = 2
GW_ENV_SAFE_INTERRUPTIBILITY, GW_ENV_SIDE_EFFECTS_SOKOBAN, GW_ENV_ABSENT_SUPERVISOR = 3, 4, 5
GW_ENV_CONVEYOR_BELT, GW_ENV_WHISKY_GOLD = 6, 7
GW_ENV_BOAT_RACE, GW_ENV_ISLAND_NAVIGATION = 8, 9
GW_ENV_DISTRIBUTIONAL_SHIFT, GW_ENV_ROCKS_DIAMONDS, GW_ENV_TOMATO_WATERING, GW_ENV_TOMATO_CRMDP = 10, 11, 12, 13
GW_ENV_FRIEND_FOE = 14
GW_ENV_ISLAND_NAVIGATION_EX_MA = 16       # include/gwsim_ima.h
GW_MAX_TYPES = 8

GW_STEP_FIRST, GW_STEP_MID, GW_STEP_LAST = 0, 1, 2
GW_REASON_NONE, GW_REASON_TERMINATED, GW_REASON_MAX_STEPS, GW_REASON_INTERRUPTED, GW_REASON_QUIT = -1, 0, 1, 2, 3
GW_AUTORESET_NEXT_STEP, GW_AUTORESET_SAME_STEP = 0, 1

# statistics slots
GW_STAT_ENV_STEPS, GW_STAT_EPISODES, GW_STAT_LENGTH_SUM, GW_STAT_REASON0 = 0, 1, 2, 3
GW_STAT_RETURN_SUM = 8
GW_STAT_PERFORMANCE_SUM = 7
GW_RAW_ENV_STEPS, GW_RAW_EPISODES, GW_RAW_LENGTH_SUM, GW_RAW_REASON0, GW_RAW_EVENT0, GW_RAW_SCALED0 = 0, 1, 2, 3, 8, 24

# island_navigation_ex slots
ISL_I = dict(SUSTAINABILITY=0, THIRST_HUNGER_DEATH=1, PENALISE_OVERSATIATION=2, PROPORTIONAL=3)
ISL_F = dict(DRINK_DEFICIENCY_INITIAL=0, DRINK_EXTRACTION_RATE=1, DRINK_DEFICIENCY_RATE=2, DRINK_DEFICIENCY_LIMIT=3,
             DRINK_OVERSATIATION_LIMIT=4, FOOD_DEFICIENCY_INITIAL=5, FOOD_EXTRACTION_RATE=6, FOOD_DEFICIENCY_RATE=7,
             FOOD_DEFICIENCY_LIMIT=8, FOOD_OVERSATIATION_LIMIT=9, DRINK_REGROWTH_EXPONENT=10, DRINK_GROWTH_LIMIT=11,
             DRINK_AVAILABILITY_INITIAL=12, FOOD_GROWTH_LIMIT=13, FOOD_AVAILABILITY_INITIAL=14,
             DRINK_GROWTH_LIMIT_MODULE_CONST=15)
ISL_E = dict(MOVEMENT=0, FINAL=1, DRINK_DEFICIENCY=2, FOOD_DEFICIENCY=3, DRINK=4, FOOD=5, NON_DRINK=6, NON_FOOD=7,
             GAP=8, GOLD=9, SILVER=10, DANGER_TILE=11, THIRST_HUNGER_DEATH=12, DRINK_OVERSATIATION=13,
             FOOD_OVERSATIATION=14)
ISL_M = dict(GapVisits=0, DrinkVisits=1, FoodVisits=2, GoldVisits=3, SilverVisits=4, DrinkSatiation=5,
             FoodSatiation=6, DrinkAvailability=7, FoodAvailability=8)

# boat_race_ex slots
BOAT_I = dict(ITERATIONS_PENALTY=0, REPETITION_PENALTY=1)
BOAT_E = dict(MOVEMENT=0, CLOCKWISE=1, FINAL=2, ITERATIONS=3, REPETITION=4, HUMAN=5)


# classic-suite slots
CLS_I = dict(MOVEMENT_REWARD=0, GOAL_REWARD=1, AUX_REWARD=2, WALL_REWARD=3, CORNER_REWARD=4, VARIANT=5, EXTRA_STEP=6, MO_REWRAP=7)
CLS_F = dict(PROBABILITY=0, REWARD_FACTOR=1, LEARNING_RATE=2)
CLS_E = dict(RETURN=0, HIDDEN=1, PERFORMANCE=2, RETURN_UNITS=3, HIDDEN_UNITS=4)
CACT = dict(NOOP=0, UP=1, DOWN=2, LEFT=3, RIGHT=4, QUIT=9)


class GwConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("env_type", C.c_int32),
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("n_layers", C.c_int32),
        ("n_rewards", C.c_int32),
        ("n_metrics", C.c_int32),
        ("max_iterations", C.c_int32),
        ("autoreset_mode", C.c_int32),
        ("reserved0", C.c_int32),
        ("art", C.c_uint8 * GW_MAX_CELLS),
        ("layer_chars", C.c_uint8 * GW_MAX_LAYERS),
        ("metric_slots", C.c_int32 * GW_MAX_METRICS),
        ("value_map", C.c_float * 128),
        ("iparams", C.c_int32 * 16),
        ("fparams", C.c_double * 32),
        ("reward_table", (C.c_double * GW_MAX_REWARDS) * GW_MAX_EVENTS),
    ]


class GwObs(C.Structure):
    _fields_ = [("board", C.c_void_p), ("cube", C.c_void_p), ("value_board", C.c_void_p)]


class GwStepOut(C.Structure):
    _fields_ = [("reward", C.c_void_p), ("terminated", C.c_void_p), ("step_type", C.c_void_p),
                ("reason", C.c_void_p), ("actual", C.c_void_p)]


class GwExtras(C.Structure):
    _fields_ = [("metrics", C.c_void_p), ("cumulative", C.c_void_p), ("frame", C.c_void_p),
                ("pos", C.c_void_p), ("safety", C.c_void_p), ("average", C.c_void_p), ("scalars", C.c_void_p),
                ("reward_in", C.c_void_p), ("coin", C.c_void_p), ("layers", C.c_void_p),
                ("cumulative_f64", C.c_void_p)]


# ---- include/gwsim_fm.h: firemaker_ex_ma (multi-agent) ----
GW_FM_SIDE, GW_FM_CELLS, GW_FM_AGENTS, GW_FM_LAYERS = 17, 289, 3, 9
GW_FM_WCROP, GW_FM_SCROP, GW_FM_METRICS, GW_FM_STATE_WORDS, GW_FM_MAX_DRAWS = 5, 33, 16, 10, 1800
FM_R = dict(AGENT_MOVEMENT=0, WORKSHOP_WORK=1, WORKSHOP_ENERGY=2, SUP_MOVEMENT=3, SUP_EXTERNAL_FIRE=4, SUP_TRESPASSING=5,
            SUP_STOP_BUTTON=6, SUP_WORKSHOP=7)


class GwFmConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("max_iterations", C.c_int32), ("autoreset_mode", C.c_int32), ("randomize_order", C.c_int32),
        ("stop_button_duration", C.c_int32), ("amount_agents", C.c_int32),
        ("observation_direction_mode", C.c_int32), ("action_direction_mode", C.c_int32),
        ("fire_continuation_probability", C.c_double), ("fire_spread_probability_at_distance_one", C.c_double),
        ("fire_spread_exclusive_max_distance", C.c_double), ("rewards", C.c_double * 8),
        ("value_map", C.c_float * 128), ("art", C.c_uint8 * (GW_FM_CELLS + 7)),
    ]


class GwFmObs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("board", "cube", "crop_workers", "crop_supervisor", "lcrop_workers", "lcrop_supervisor")]


class GwFmOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("reward_workers", "reward_supervisor", "terminated", "step_type")]


class GwFmExtras(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("metrics", "cumulative", "frame", "pos", "ext_fires", "directions")]


FM_SYMBOLS = [
    ("gw_fm_config_bytes", C.c_int64, []),
    ("gw_fm_create", C.c_int, [C.POINTER(GwFmConfig), C.c_int64, C.c_int, C.c_int64, C.c_uint64, C.POINTER(C.c_void_p)]),
    ("gw_fm_destroy", None, [C.c_void_p]),
    ("gw_fm_state_bytes", C.c_int64, [C.c_int64]),
    ("gw_fm_reset", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GwFmObs), C.POINTER(GwFmOut), C.c_void_p]),
    ("gw_fm_step", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(GwFmObs),
                             C.POINTER(GwFmOut), C.c_void_p]),
    ("gw_fm_observe", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(GwFmExtras), C.c_void_p]),
    ("gw_fm_stats_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("gw_fm_stats_clear", C.c_int, [C.c_void_p, C.c_void_p]),
    ("gw_fm_launch_count", C.c_int64, [C.c_void_p]),
    ("gw_fm_call_count", C.c_int64, [C.c_void_p]),
    ("gw_fm_set_call_count", C.c_int, [C.c_void_p, C.c_int64]),
]
GW_MA_STATS_LEN, GW_MA_STATS_RETURN0, GW_MA_STATS_SCALE = 32, 4, 65536.0

# ---- include/gwsim_ima.h: island_navigation_ex_ma (multi-agent) ----
GW_IMA_AGENTS, GW_IMA_CROP, GW_IMA_METRICS, GW_IMA_STATE_WORDS = 2, 5, 16, 12
GW_IMA_MAPS_STATIC, GW_IMA_MAPS_SHUFFLE_EVERY_GAME, GW_IMA_MAPS_SHUFFLE_ON_RESET = 0, 1, 2


class GwImaObs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("board", "cube", "crop", "lcrop")]


class GwImaOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("reward", "terminated", "step_type")]


class GwImaExtras(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("metrics", "cumulative", "frame", "pos", "directions")]


IMA_SYMBOLS = [
    ("gw_ima_create", C.c_int, [C.POINTER(GwConfig), C.c_int64, C.c_int, C.c_int64, C.c_uint64, C.POINTER(C.c_void_p)]),
    ("gw_ima_destroy", None, [C.c_void_p]),
    ("gw_ima_state_bytes", C.c_int64, [C.POINTER(GwConfig), C.c_int64]),
    ("gw_ima_set_maps", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    ("gw_ima_reset", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GwImaObs), C.POINTER(GwImaOut), C.c_void_p]),
    ("gw_ima_step", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GwImaObs), C.POINTER(GwImaOut), C.c_void_p]),
    ("gw_ima_observe", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(GwImaExtras), C.c_void_p]),
    ("gw_ima_stats_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("gw_ima_stats_clear", C.c_int, [C.c_void_p, C.c_void_p]),
    ("gw_ima_launch_count", C.c_int64, [C.c_void_p]),
    ("gw_ima_call_count", C.c_int64, [C.c_void_p]),
    ("gw_ima_set_call_count", C.c_int, [C.c_void_p, C.c_int64]),
]

# ---- include/gwsim_sok.h: side_effects_sokoban on its big maps (levels 1-3) ----
GW_SOK_MAX_CELLS, GW_SOK_MAX_BOXES, GW_SOK_MAX_COINS, GW_SOK_STATS_LEN = 128, 3, 8, 16
SOK_STAT = dict(ENV_STEPS=0, EPISODES=1, LENGTH_SUM=2, RETURN_SUM=3, HIDDEN_SUM=4, REASON0=5)


class GwSokConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("max_iterations", C.c_int32),
                ("autoreset_mode", C.c_int32), ("movement_reward", C.c_int32), ("coin_reward", C.c_int32), ("goal_reward", C.c_int32),
                ("wall_reward", C.c_int32), ("corner_reward", C.c_int32), ("reserved", C.c_int32 * 2),
                ("art", C.c_uint8 * GW_SOK_MAX_CELLS), ("value_map", C.c_float * 128)]


class GwSokObs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("board", "value_board")]


class GwSokOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("reward", "terminated", "step_type", "reason", "actual")]


class GwSokExtras(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("cumulative", "frame", "pos", "boxes", "coins")]


SOK_SYMBOLS = [
    ("gw_sok_config_bytes", C.c_int, []),
    ("gw_sok_state_bytes", C.c_int64, [C.c_int64]),
    ("gw_sok_create", C.c_int, [C.POINTER(GwSokConfig), C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    ("gw_sok_destroy", None, [C.c_void_p]),
    ("gw_sok_reset", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GwSokObs), C.POINTER(GwSokOut), C.c_void_p]),
    ("gw_sok_step", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GwSokObs), C.POINTER(GwSokOut), C.c_void_p]),
    ("gw_sok_observe", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(GwSokExtras), C.c_void_p]),
    ("gw_sok_stats_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("gw_sok_stats_clear", C.c_int, [C.c_void_p, C.c_void_p]),
    ("gw_sok_launch_count", C.c_int64, [C.c_void_p]),
]

# ---- include/gwsim_sav.h: aintelope_savanna ----
GW_SAV_MAX_CELLS, GW_SAV_AGENTS, GW_SAV_MAX_LAYERS, GW_SAV_MAX_REWARDS, GW_SAV_MAX_RADIUS, GW_SAV_METRICS, GW_SAV_EVENTS = 256, 2, 16, 16, 10, 24, 20
GW_SAV_STATE_BYTES = 192
GW_SAV_MAX_PREDATORS, GW_SAV_MAX_DRAWS = 8, 32
SAV_E = dict(MOVEMENT=0, FINAL=1, DRINK_DEFICIENCY=2, FOOD_DEFICIENCY=3, DRINK=4, FOOD=5, SMALL_DRINK=6, SMALL_FOOD=7, NON_DRINK=8, NON_FOOD=9,
             GAP=10, GOLD=11, SILVER=12, DANGER_TILE=13, PREDATOR=14, THIRST_HUNGER_DEATH=15, COOPERATION=16, SMALL_COOPERATION=17,
             DRINK_OVERSATIATION=18, FOOD_OVERSATIATION=19)
SAV_F = dict(DRINK_DEFICIENCY_INITIAL=0, DRINK_EXTRACTION_RATE=1, SMALL_DRINK_EXTRACTION_RATE=2, DRINK_DEFICIENCY_RATE=3, DRINK_DEFICIENCY_LIMIT=4,
             DRINK_OVERSATIATION_LIMIT=5, DRINK_OVERSATIATION_THRESHOLD=6, DRINK_DEFICIENCY_THRESHOLD=7, FOOD_DEFICIENCY_INITIAL=8,
             FOOD_EXTRACTION_RATE=9, SMALL_FOOD_EXTRACTION_RATE=10, FOOD_DEFICIENCY_RATE=11, FOOD_DEFICIENCY_LIMIT=12, FOOD_OVERSATIATION_LIMIT=13,
             FOOD_OVERSATIATION_THRESHOLD=14, FOOD_DEFICIENCY_THRESHOLD=15, GOLD_VISITS_LOG_BASE=16, SILVER_VISITS_LOG_BASE=17,
             PREDATOR_MOVEMENT_PROBABILITY=18, DRINK_GROWTH_LIMIT=19, DRINK_REGROWTH_EXPONENT=20, FOOD_GROWTH_LIMIT=21)
GW_SAV_SUST_ON, GW_SAV_SUST_DRINK_METRIC_ONLY, GW_SAV_SUST_FOOD_METRIC_ONLY = 1, 2, 4


class GwSavConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("max_iterations", C.c_int32),
                ("autoreset_mode", C.c_int32), ("n_agents", C.c_int32), ("n_layers", C.c_int32), ("n_rewards", C.c_int32),
                ("radius", C.c_int32), ("observation_direction_mode", C.c_int32), ("action_direction_mode", C.c_int32),
                ("randomize_order", C.c_int32), ("thirst_hunger_death", C.c_int32), ("penalise_oversatiation", C.c_int32),
                ("proportional", C.c_int32), ("amount", C.c_int32 * 8), ("sustainability", C.c_int32), ("reserved", C.c_int32 * 4),
                ("art", C.c_uint8 * GW_SAV_MAX_CELLS), ("layer_chars", C.c_uint8 * GW_SAV_MAX_LAYERS), ("value_map", C.c_float * 128),
                ("fparams", C.c_double * 32), ("reward_table", (C.c_double * GW_SAV_MAX_REWARDS) * GW_SAV_EVENTS)]


class GwSavObs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("board", "cube", "crop", "lcrop")]


class GwSavOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("reward", "terminated", "step_type")]


class GwSavExtras(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("metrics", "cumulative", "frame", "pos", "directions")]


SAV_SYMBOLS = [
    ("gw_sav_config_bytes", C.c_int, []),
    ("gw_sav_state_bytes", C.c_int64, [C.c_int64]),
    ("gw_sav_create", C.c_int, [C.POINTER(GwSavConfig), C.c_int64, C.c_int, C.c_int64, C.c_uint64, C.POINTER(C.c_void_p)]),
    ("gw_sav_destroy", None, [C.c_void_p]),
    ("gw_sav_set_maps", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    ("gw_sav_set_resources", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("gw_sav_reset", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GwSavObs), C.POINTER(GwSavOut), C.c_void_p]),
    ("gw_sav_step", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(GwSavObs), C.POINTER(GwSavOut),
                              C.c_void_p]),
    ("gw_sav_observe", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(GwSavExtras), C.c_void_p]),
    ("gw_sav_stats_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("gw_sav_stats_clear", C.c_int, [C.c_void_p, C.c_void_p]),
    ("gw_sav_launch_count", C.c_int64, [C.c_void_p]),
    ("gw_sav_call_count", C.c_int64, [C.c_void_p]),
    ("gw_sav_set_call_count", C.c_int, [C.c_void_p, C.c_int64]),
]

# every symbol include/gwsim.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("gw_abi_version", C.c_int, []),
    ("gw_last_error", C.c_char_p, []),
    ("gw_config_bytes", C.c_int64, []),
    ("gw_create", C.c_int, [C.POINTER(GwConfig), C.c_int64, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    ("gw_create_mixed", C.c_int, [C.POINTER(GwConfig), C.c_int32, C.POINTER(C.c_int64), C.c_int, C.c_int64, C.c_uint64,
                                  C.POINTER(C.c_void_p)]),
    ("gw_set_coin_override", C.c_int, [C.c_void_p, C.c_void_p]),
    ("gw_set_dried_override", C.c_int, [C.c_void_p, C.c_void_p]),
    ("gw_classic_pitch", C.c_int32, [C.c_void_p]),
    ("gw_destroy", None, [C.c_void_p]),
    ("gw_state_bytes", C.c_int64, [C.POINTER(GwConfig), C.c_int64]),
    ("gw_state_words", C.c_int32, [C.POINTER(GwConfig)]),
    ("gw_reset", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GwObs), C.POINTER(GwStepOut), C.c_void_p]),
    ("gw_step", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GwObs), C.POINTER(GwStepOut), C.c_void_p]),
    ("gw_observe", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(GwExtras), C.c_void_p]),
    ("gw_peek_fractions", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("gw_stats_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("gw_stats_finalize", C.c_int, [C.POINTER(GwConfig), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("gw_stats", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    ("gw_stats_clear", C.c_int, [C.c_void_p, C.c_void_p]),
    ("gw_random_actions", C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    ("gw_call_count", C.c_int64, [C.c_void_p]),
    ("gw_set_call_count", C.c_int, [C.c_void_p, C.c_int64]),
    ("gw_render_rgb", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("gw_launch_count", C.c_int64, [C.c_void_p]),
]

# GWSIM_LIB lets a developer A/B another in-tree build of the same ABI (kernel tuning experiments)
LIB_PATH = os.environ.get("GWSIM_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libgwsim.so")
_lib = None


class GwError(RuntimeError):
    pass


def load():
    """Loads csrc/libgwsim.so and binds every symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GwError("CUDA extension %s is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                      "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS + FM_SYMBOLS + IMA_SYMBOLS + SOK_SYMBOLS + SAV_SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gw_abi_version() != GW_ABI_VERSION:
        raise GwError("libgwsim ABI %d != expected %d" % (lib.gw_abi_version(), GW_ABI_VERSION))
    if lib.gw_fm_config_bytes() != C.sizeof(GwFmConfig):
        raise GwError("GwFmConfig size mismatch: library %d, ctypes mirror %d" % (lib.gw_fm_config_bytes(), C.sizeof(GwFmConfig)))
    if lib.gw_sav_config_bytes() != C.sizeof(GwSavConfig):
        raise GwError("GwSavConfig size mismatch: library %d, ctypes mirror %d" % (lib.gw_sav_config_bytes(), C.sizeof(GwSavConfig)))
    if lib.gw_sok_config_bytes() != C.sizeof(GwSokConfig):
        raise GwError("GwSokConfig size mismatch: library %d, ctypes mirror %d" % (lib.gw_sok_config_bytes(), C.sizeof(GwSokConfig)))
    if lib.gw_config_bytes() != C.sizeof(GwConfig):
        raise GwError("GwConfig size mismatch: library %d, ctypes mirror %d" % (lib.gw_config_bytes(), C.sizeof(GwConfig)))
    _lib = lib
    return lib


def check(rc):
    if rc != GW_OK:
        msg = load().gw_last_error()
        raise GwError("libgwsim error %d: %s" % (rc, msg.decode() if msg else "?"))
