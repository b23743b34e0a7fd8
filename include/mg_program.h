/*
 * mg_program.h -- layout of the COMPILED GAME PROGRAM consumed by the step kernels.
 *
 * The reference builds opaque C++ Handler/Filter/Mutation objects out of the Pydantic config
 * (python/src/mettagrid/config/mettagrid_c_config.py:576-1007).  Here the same config is
 * lowered once on the host (mettagrid_b200/compiler.py) into one flat table of int32 words:
 * a header, then sections.  The same blob is interpreted by
 *   - the sm_100a step kernel        (mettagrid_b200/csrc/step_kernel.cu)
 *   - the CPU oracle restatement     (oracle/mg_oracle.cpp; test infrastructure only)
 * This file is a DATA FORMAT: it holds no algorithm.
 *
 * All offsets are in 32-bit words from the start of the blob.  Floats are stored as their
 * IEEE-754 bit patterns.  "list" = (offset into MGS_POOL, count).
 */
#ifndef MG_PROGRAM_H_
#define MG_PROGRAM_H_

#include <stdint.h>

#define MG_MAGIC 0x4D47B200
#define MG_VERSION 8

/* ---- header word indices ------------------------------------------------------------ */
enum {
  MGH_MAGIC = 0,
  MGH_VERSION,
  MGH_TOTAL_WORDS,
  MGH_H,               /* map height (rows) */
  MGH_W,               /* map width (cols) */
  MGH_NUM_AGENTS,      /* agents per env (A) */
  MGH_NUM_TOKENS,      /* observation token budget (T) */
  MGH_NUM_RESOURCES,   /* R (<= 13) */
  MGH_NUM_TAGS,
  MGH_TAG_WORDS,       /* TW = ceil(num_tags/32), >= 1 */
  MGH_TOKEN_BASE,      /* token_value_base */
  MGH_INV_DIGITS,      /* tokens needed for 65535 in that base */
  MGH_MAX_STEPS,
  MGH_EPISODE_TRUNCATES,
  MGH_OBS_H,
  MGH_OBS_W,
  MGH_NUM_OFFSETS,     /* cells in the observation shape */
  MGH_GLOBAL_FLAGS,    /* MGG_* bits */
  MGH_FEAT_GROUP,
  MGH_FEAT_EPISODE_PCT,
  MGH_FEAT_LAST_ACTION,
  MGH_FEAT_LAST_REWARD,
  MGH_FEAT_VIBE,
  MGH_FEAT_TAG,
  MGH_FEAT_LP_EAST,
  MGH_FEAT_LP_WEST,
  MGH_FEAT_LP_NORTH,
  MGH_FEAT_LP_SOUTH,
  MGH_FEAT_AGENT_ID,
  MGH_FEAT_AOE_MASK,          /* 0 = disabled */
  MGH_FEAT_LAST_ACTION_MOVE,  /* 0 = disabled */
  MGH_NUM_ACTIONS,
  MGH_MAX_PRIORITY,
  MGH_PRIORITY_MASK,   /* bit p set when some action has priority p */
  MGH_NUM_AGENT_STATS, /* S_A */
  MGH_NUM_GAME_STATS,  /* S_G */
  MGH_MAX_OBJECTS,     /* object pool capacity per env (slot 0 unused) */
  MGH_NUM_TEMPLATES,
  MGH_OBJ_STRIDE,      /* words per object record */
  MGH_AGENT_STRIDE,    /* words per agent record */
  MGH_COVER_WORDS,     /* ceil(H*W/32) */
  MGH_MAX_REWARDS,     /* max reward entries over agents */
  MGH_HP_RESOURCE,     /* resource id named "hp" or -1 (objects/agent.cpp:117) */
  /* well-known agent stat ids */
  MGH_ST_ACTION_FAILED,
  MGH_ST_INVALID_INDEX,
  MGH_ST_MAX_SWM,
  MGH_ST_CELL_VISITED,
  MGH_ST_UNIQUE_VISITED,
  MGH_ST_MAX_DIST,
  MGH_ST_DEATH,
  MGH_ST_ACTIONS_SWAP,
  MGH_ST_NOOP_SUCCESS,
  MGH_ST_NOOP_FAILED,
  MGH_ST_MOVE_SUCCESS,
  MGH_ST_MOVE_FAILED,
  MGH_ST_VIBE_SUCCESS,
  MGH_ST_VIBE_FAILED,
  /* well-known game stat ids */
  MGH_GST_TOKENS_WRITTEN,
  MGH_GST_TOKENS_DROPPED,
  MGH_GST_TOKENS_FREE,
  /* counts */
  MGH_NUM_MOVE_HANDLERS,
  MGH_NUM_OBS_VALUES,
  MGH_NUM_EVENTS_SCHED,
  MGH_NUM_TERRITORIES,
  MGH_NUM_MQ,           /* materialized queries */
  MGH_GAME_ON_TICK,     /* handler id or -1 */
  MGH_NUM_DYN_TAGS,     /* tags that can be added at run time (need insertion stamps) */
  MGH_MAX_AOE_SOURCES,  /* capacity of the per-env AOE source table */
  MGH_MAX_TERR_SOURCES,
  MGH_PROXY_TEMPLATE,   /* inert template (kind 3) used by territory proxy-cell objects */
  MGH_SPAWN_AOES,       /* max AOE configs on any template a spawn mutation can create */
  MGH_TOK_CAP,          /* most observation tokens any object can emit (per-object token cache size) */
  /* compile-time effect analysis (mettagrid_b200/compiler.py: analyze_effects) */
  MGH_MOVE_REACH,       /* longest line a move handler scans (max max_range over the chain) */
  MGH_CHAIN_EMPTY_CLASS,/* MGC_* class of the custom move handlers that accept an empty target cell */
  MGH_PAR_FLAGS,        /* MGP_* bits: which per-agent phases may run one lane per agent */
  MGH_SPAWN_CLASS,      /* highest MGC_* class of any template a SpawnObject mutation can create (0 = nothing spawns) */
  /* section offsets */
  MGS_OFFSETS,      /* NUM_OFFSETS x (dr, dc) */
  MGS_ACTIONS,      /* NUM_ACTIONS x MG_ACTION_WORDS */
  MGS_MOVE_CHAIN,   /* NUM_MOVE_HANDLERS x MG_MOVEH_WORDS */
  MGS_HANDLERS,     /* n x MG_HANDLER_WORDS */
  MGS_FILTERS,      /* n x MG_FILTER_WORDS */
  MGS_MUTATIONS,    /* n x MG_MUTATION_WORDS */
  MGS_VALUES,       /* n x MG_VALUE_WORDS (game-value expression nodes) */
  MGS_QUERIES,      /* n x MG_QUERY_WORDS */
  MGS_TEMPLATES,    /* NUM_TEMPLATES x MG_TEMPLATE_WORDS */
  MGS_LIMITS,       /* n x MG_LIMIT_WORDS */
  MGS_AOES,         /* n x MG_AOE_WORDS */
  MGS_EVENTS,       /* n x MG_EVENT_WORDS */
  MGS_SCHEDULE,     /* NUM_EVENTS_SCHED x (timestep, event id) */
  MGS_TERRITORIES,  /* NUM_TERRITORIES x MG_TERR_WORDS */
  MGS_MQ,           /* NUM_MQ x (tag id, query id) */
  MGS_OBS_VALUES,   /* NUM_OBS_VALUES x (feature id, value node) */
  MGS_RES_STATS,    /* R x 4: agent stat ids gained, lost, amount, deposited */
  MGS_RES_GSTATS,   /* R x 1: game stat id of "<res>.amount" (InventoryValue w/o actor) or -1 */
  MGS_INV_FEATS,    /* R x INV_DIGITS feature ids */
  MGS_DYN_TAGS,     /* NUM_TAGS x 1: dynamic-tag slot or -1 */
  MGS_TMPL_CLASS,   /* NUM_TEMPLATES x 1: MGC_* class (bits 0-1) | MGC_DISPLACES of a move whose line holds such an object */
  MGS_POOL,         /* variable-length int lists */
  MGH_HEADER_WORDS
};

/* Effect classes of one agent's action (what it may read / write besides itself and the objects in its line):
   LOCAL  -- only the actor, the objects in the scanned line and those cells;
   SHARED -- also env-wide structures whose update order matters (game stats, tag index, object table, AOE tables):
             such actions run in the shuffled order relative to each other, but independently of LOCAL ones elsewhere;
   SERIAL -- reads or writes arbitrary objects (queries, raycasts, push chains): the whole pass runs in order. */
enum { MGC_LOCAL = 0, MGC_SHARED = 1, MGC_SERIAL = 2 };
#define MGC_DISPLACES 4 /* may move the TARGET agent (swap): the target's own move then starts from the actor's cell */
/* MGH_PAR_FLAGS */
#define MGP_ON_TICK 1 /* agent on_tick handlers only read and write their own agent */
#define MGP_AOE 2     /* AOE / territory handlers only write their target agent and read nothing another lane writes */
#define MGP_ACTIONS 4 /* action passes may resolve independent agents concurrently */

/* MGH_GLOBAL_FLAGS bits (cpp/bindings/mettagrid_c.cpp:700-742) */
#define MGG_EPISODE_PCT 1
#define MGG_LAST_ACTION 2
#define MGG_LAST_ACTION_MOVE 4
#define MGG_LAST_REWARD 8
#define MGG_LOCAL_POSITION 16

/* ---- actions (actions/action_handler_factory.cpp:15-78) ------------------------------ */
#define MG_ACTION_WORDS 4 /* kind, arg, priority, is_vibe */
enum { MGA_NOOP = 0, MGA_MOVE = 1, MGA_CHANGE_VIBE = 2 };

#define MG_MOVEH_WORDS 4 /* handler id, max_range, accepts_empty (actions/move.hpp:26-40), builtin */
/* builtin: the two handlers the reference always appends (action_handler_factory.cpp:33-45) are
   also emitted as ordinary handlers; the marker lets an engine take a table-free fast path. */
enum { MGMB_GENERIC = 0, MGMB_RELOCATE = 1, MGMB_USE_TARGET = 2 };

/* ---- handlers (handler/handler.cpp:76-103, multi_handler.cpp:8-21) -------------------- */
#define MG_HANDLER_WORDS 5 /* kind, a, b, c, d */
enum {
  MGHK_SIMPLE = 0,      /* a=first filter, b=#filters, c=first mutation, d=#mutations */
  MGHK_FIRST_MATCH = 1, /* a=pool offset of child handler ids, b=#children */
  MGHK_ALL = 2
};

enum { MGE_ACTOR = 0, MGE_TARGET = 1, MGE_SOURCE = 2 };

/* ---- filters (handler/filters/ *.hpp) -------------------------------------------------- */
#define MG_FILTER_WORDS 6 /* op, entity, a, b, c, d */
enum {
  MGF_VIBE = 0,          /* a=vibe id */
  MGF_RESOURCE,          /* a=resource, b=min amount */
  MGF_SHARED_TAG_PREFIX, /* a=pool offset of TW mask words */
  MGF_TAG_PREFIX,        /* a=pool offset of TW mask words */
  MGF_GAME_VALUE,        /* a=value node, b=threshold node */
  MGF_NEG,               /* a=first child filter, b=#children : NOT(AND(children)) */
  MGF_OR,                /* a=first child filter, b=#children */
  MGF_MAX_DISTANCE,      /* a=query id or -1, b=radius */
  MGF_QUERY_RESOURCE,    /* a=query id, b=pool offset of (resource, min) pairs, c=#pairs */
  MGF_TARGET_LOC_EMPTY,
  MGF_TARGET_IS_USABLE,
  MGF_PERIODIC           /* a=period, b=start_on */
};

/* ---- mutations (handler/mutations/ *.hpp) ---------------------------------------------- */
#define MG_MUTATION_WORDS 8 /* op, e1, e2, a, b, c, d, e */
enum {
  MGM_RESOURCE_DELTA = 0, /* e1=entity, a=resource, b=delta */
  MGM_RESOURCE_TRANSFER,  /* e1=source, e2=dest, a=resource, b=amount(-1 all), c=remove_source_when_empty */
  MGM_CLEAR_INVENTORY,    /* e1=entity, a=pool offset of resources, b=count (0 = everything) */
  MGM_ATTACK,             /* a=weapon, b=armor, c=health, d=damage pct */
  MGM_STATS,              /* a=stat id, b=0 game/1 agent, c=0 target/1 actor, d=value node; e=1 when the value is
                             "this stat + integer constant" (logStat), the constant's float bits then sit in e2 */
  MGM_ADD_TAG,            /* e1, a=tag */
  MGM_REMOVE_TAG,         /* e1, a=tag */
  MGM_GAME_VALUE,         /* e1=target entity, a=value node, b=source node */
  MGM_RECOMPUTE_MQ,       /* a=tag id */
  MGM_QUERY_INVENTORY,    /* e1=source entity, a=query, b=pool off of (res, delta) pairs, c=#pairs,
                             d=has_source, e=pool off of R game-stat ids (or -1) */
  MGM_REMOVE_TAGS_PREFIX, /* e1, a=pool offset of tag ids, b=count */
  MGM_RELOCATE,
  MGM_SWAP,
  MGM_USE_TARGET,
  MGM_SPAWN_OBJECT,       /* a=template or -1 */
  MGM_RAYCAST_SPAWN,      /* a=template or -1, b=pool off (dr,dc) pairs, c=#dirs, d=range node,
                             e=first blocker filter, e2=#blockers */
  MGM_CHANGE_VIBE,        /* e1, a=vibe */
  MGM_PUSH_OBJECT
};

/* ---- game values (core/game_value.cpp:14-148) ------------------------------------------ */
#define MG_VALUE_WORDS 6 /* op, scope, a, b, c, d */
enum { MGSC_AGENT = 0, MGSC_GAME = 1 };
enum {
  MGV_INVENTORY = 0,   /* a=resource */
  MGV_STAT,            /* a=stat id in the scope's table, b=delta flag */
  MGV_CONST,           /* a=float bits */
  MGV_QUERY_INVENTORY, /* a=resource, b=query */
  MGV_QUERY_COUNT,     /* b=query */
  MGV_SUM,             /* a=pool off of child nodes, b=n, c=pool off of weights (float bits) or -1, d=log */
  MGV_RATIO,           /* a=numerator node, b=denominator node */
  MGV_MAX,             /* a=pool off, b=n */
  MGV_MIN
};

/* ---- queries (core/query_system.cpp:178-330) ------------------------------------------- */
#define MG_QUERY_WORDS 12
/* kind, max_items node (-1 = unlimited), order_by_random, a..: see below */
enum {
  MGQ_TAG = 0,  /* [3]=tag id, [4]=first filter, [5]=#filters */
  MGQ_FILTERED, /* [3]=source query, [4]=first filter, [5]=#filters */
  MGQ_CLOSURE,  /* [3]=source, [4]=candidates (-1 none), [5]=first edge filter, [6]=#edge,
                   [7]=first result filter, [8]=#result */
  MGQ_RAYCAST   /* [3]=source, [4]=range node, [5]=pool off (dr,dc), [6]=#dirs, [7]=first blocker,
                   [8]=#blockers, [9]=include_blocker */
};

/* ---- object templates (one per map-cell name; core/grid_object_factory.cpp:62-104) ------ */
#define MG_TEMPLATE_WORDS 24
enum {
  MGT_KIND = 0,     /* 0 wall, 1 agent, 2 object, 3 territory proxy cell */
  MGT_TYPE_ID,
  MGT_VIBE,
  MGT_TAGS,         /* pool offset of TW words */
  MGT_INIT_INV,     /* pool offset of (resource, amount) pairs, in emission (dict) order */
  MGT_INIT_INV_N,
  MGT_LIMIT_OF,     /* pool offset of R words: limit id (global index into MGS_LIMITS) or -1 */
  MGT_LIMIT_ORDER,  /* pool offset: limit ids in enforce order (objects/inventory.cpp:143-149) */
  MGT_LIMIT_ORDER_N,
  MGT_RES_ORDER,    /* pool offset of resources in _limits iteration order (inventory.cpp:155) */
  MGT_RES_ORDER_N,
  MGT_MODIFIER_MASK,/* bit r set if resource r is a modifier of a reachable limit */
  MGT_ON_USE,       /* handler id or -1 */
  MGT_ON_TICK,      /* agents */
  MGT_ON_AFTER_USE, /* agents */
  MGT_GROUP,        /* agents */
  MGT_REWARDS,      /* pool offset of (value node, accumulate) pairs */
  MGT_REWARDS_N,
  MGT_AOES,         /* first AOE config index */
  MGT_AOES_N,
  MGT_TERR,         /* pool offset of (territory, strength, decay) */
  MGT_TERR_N,
  MGT_TAG_REMOVE,   /* pool offset of (tag, handler id) pairs, in per-tag list order */
  MGT_TAG_REMOVE_N
};

#define MG_LIMIT_WORDS 6 /* min, max, pool off modifiers (item,bonus), #modifiers, pool off members, #members */

#define MG_AOE_WORDS 10
/* radius, is_static, effect_self, first filter, #filters, first mutation, #mutations,
   pool off presence deltas (res, delta), #deltas, reserved */

#define MG_EVENT_WORDS 8
/* query, max_targets, first filter, #filters, first mutation, #mutations, fallback event or -1, reserved */

#define MG_TERR_WORDS 8
/* pool off prefix tag ids, #ids, pool off on_enter handler ids, n, on_exit off, n, presence off, n */

/* ---- per-env state layout --------------------------------------------------------------- */
/* object record words (stride = MGH_OBJ_STRIDE, multiple of 4 words) */
enum {
  MGO_LOC = 0,  /* r << 16 | c */
  MGO_VISITED,  /* unused: GridObject::visited lives in a compact per-env array (mg_state.h: visited) */
  MGO_META,     /* template (16) | vibe (8) << 16 | flags (8) << 24 */
  MGO_AGENT,    /* agent index or -1 */
  MGO_INVORD_LO,/* inventory iteration order: 4-bit resource ids, most recent first */
  MGO_INVORD_HI,/* ... top nibble (bits 60-63) = number of present resources */
  MGO_ID,       /* GridObject::id (insertion order; query result order) */
  MGO_NTOK,     /* cached token count, MG_TOK_DIRTY when the cache must be rebuilt */
  MGO_TAGS      /* TW words, then ceil(R/2) words of u16 amounts, then ceil(TOK_CAP/2) words of cached
                   (feature | value << 8) observation tokens */
};
/* first word of the cached tokens: behind the tags and the inventory, rounded up to a 16-byte boundary */
#define MG_TOKOFF(TW, R) ((MGO_TAGS + (TW) + ((R) + 1) / 2 + 3) & ~3)
#define MG_TOK_DIRTY 0xFFFFFFFFu
#define MGOF_ALIVE 1
#define MGOF_AGENT 2
#define MGOF_OBS_INV 4 /* obs_encoder set: inventory tokens are emitted (false for spawned objects) */
#define MGOF_WALL 8
#define MGOF_TERR_SRC 16 /* registered as a territory source: moving / re-tagging it invalidates the ownership map */

/* agent record words (stride = MGH_AGENT_STRIDE) */
enum {
  MGAG_OBJ = 0,       /* object slot */
  MGAG_SPAWN,         /* r << 16 | c */
  MGAG_PREV_LOC,      /* Agent::prev_location (action_handler.hpp:83-93) */
  MGAG_STEP_LOC,      /* _prev_agent_locations (mettagrid_c.cpp:929-931) */
  MGAG_SWM,           /* steps_without_motion */
  MGAG_MAX_DIST,
  MGAG_UNIQUE,        /* |unique_cells_visited| */
  MGAG_EPISODE_REWARD,/* float bits */
  MGAG_REWARD_PREV    /* MAX_REWARDS float bit patterns */
};

/* per-env scalar words */
enum {
  MGEV_STEP = 0,
  MGEV_RNG_IDX,      /* 0..624 (624 = state exhausted, next draw regenerates from 0) */
  MGEV_ERROR,        /* MGERR_* bits */
  MGEV_ERR_INFO,     /* agent | attempted << 16 for token overflow */
  MGEV_NEXT_OBJ,     /* next free object slot */
  MGEV_NEXT_ID,      /* next GridObject id */
  MGEV_EVENT_CURSOR,
  MGEV_TAG_SEQ,      /* dynamic-tag insertion counter */
  MGEV_NUM_AOE,      /* registered AOE sources */
  MGEV_NUM_AOE_PENDING,
  MGEV_NUM_TERR,
  MGEV_RESERVED,     /* entries in the per-tick territory source table */
  MGEV_TERR_STALE,   /* a territory source moved / changed tags since the ownership map was built */
  MGEV_PAD0,
  MGEV_PAD1,
  MGEV_PAD2,
  MGEV_WORDS
};
#define MGERR_TOKEN_OVERFLOW 1
#define MGERR_POOL_EXHAUSTED 2
#define MGERR_UNSUPPORTED 4
#define MGERR_BOUNDS 8 /* checked builds only (MG_CHECKED): a state accessor was handed an index outside its array */

#define MG_RNG_WORDS 624

#endif /* MG_PROGRAM_H_ */
