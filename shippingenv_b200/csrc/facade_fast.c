/* facade_fast.c -- the hot calls of the N = 1 drop-in Environment (shippingenv_b200/environment.py) in C.
 *
 * The reference's agents call `a = env.sample_action(); env.step(a)` in a Python loop
 * (agents/mcts.py:228-236, agents/sarsa.py:211-220).  With the step answered by the resident server kernel in ~5 us
 * (shipenv_serve_step), the facade's own Python -- encoding the action, packing the attributes into the host block,
 * the ctypes call, unpacking, repainting, rebuilding the snapshot -- had become as large as the device round trip.
 * This module does exactly that work for the COMMON shapes of a call and declines (returns NotImplemented) for
 * everything else, in which case environment.py runs its Python code, which stays the complete implementation and
 * the place where every exception of the reference is raised: unit moves / SELECT_PORT / TAKE_* with plain ints, a
 * ship position that is a list of two ints, the served path, a square map.  It computes nothing of the
 * environment: the transition is shipenv_serve_step's (the CUDA server kernel), called through the pointer that
 * bind() received.
 *
 * Built by __graft_entry__.build() (gcc, CPython C API) into shippingenv_b200/_facade_fast.so;
 * SHIPENV_NO_FASTPATH=1 keeps it out.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

typedef int (*serve_fn_t)(void *ctx, void *block, uint64_t seed, uint64_t env_id, int max_steps, uint32_t flags,
                          uint32_t word0);

static serve_fn_t g_serve = NULL;
static PyObject *g_errors = NULL;   /* status -> (exception class, message) */
static PyObject *g_check = NULL;    /* _native.check(rc): raises ShipEnvError */
static PyObject *g_moves = NULL;    /* (N, E, S, W) as the (dx, dy) tuples of ShipMove */
static long g_n_cargo = 50, g_n_fuel = 200, g_ext_cargo = 0x4000, g_ext_fuel = 0x6000;

/* interned attribute names */
static PyObject *s_port_positions, *s_destination, *s_origin, *s_ship_position, *s_cargo, *s_fuel, *s_clock, *s_word,
    *s_seed, *s_env_id, *s_host_ptr, *s_ragged, *s_sync_world, *s_h_int, *s_spec, *s_world_sig, *s_np_game,
    *s_build_state, *s_slot, *s_w_game, *s_w_ports, *s_w_fuel, *s_w_cargo, *s_port_fuel, *s_port_cargo;

/* offsets inside shipenv_host_block (include/shipenv.h) */
enum { OFF_SHIP = 0, OFF_REWARD = 16, OFF_CLOCK = 32, OFF_ACTION = 48, OFF_DONE = 64, OFF_STATUS = 76, OFF_NEXT = 84 };
enum { NONE_PORT = 255, ENT_BOAT = 5, ENT_TRAVEL = 7 };

#define DECLINE() do { Py_INCREF(Py_NotImplemented); result = Py_NotImplemented; goto done; } while (0)

static int as_u64(PyObject *o, uint64_t *out) {
  if (!o || !PyLong_Check(o)) return -1;
  unsigned long long v = PyLong_AsUnsignedLongLongMask(o);
  if (v == (unsigned long long)-1 && PyErr_Occurred()) { PyErr_Clear(); return -1; }
  *out = (uint64_t)v;
  return 0;
}

static int as_long_exact(PyObject *o, long *out) {
  if (!o || !PyLong_CheckExact(o)) return -1;
  int overflow = 0;
  long v = PyLong_AsLongAndOverflow(o, &overflow);
  if (overflow) return -1;
  if (v == -1 && PyErr_Occurred()) { PyErr_Clear(); return -1; }
  *out = v;
  return 0;
}

/* (pos[0], pos[1], origin, dest, cargo, fuel, clock, word, world_sig): what Environment._snapshot() builds */
static PyObject *snapshot(PyObject *env) {
  PyObject *pos = PyObject_GetAttr(env, s_ship_position);
  if (!pos) return NULL;
  if (!PyList_CheckExact(pos) || PyList_GET_SIZE(pos) != 2) {
    Py_DECREF(pos);
    Py_RETURN_NONE;
  }
  PyObject *names[7] = {s_origin, s_destination, s_cargo, s_fuel, s_clock, s_word, s_world_sig};
  PyObject *t = PyTuple_New(9);
  if (!t) { Py_DECREF(pos); return NULL; }
  PyObject *p0 = PyList_GET_ITEM(pos, 0), *p1 = PyList_GET_ITEM(pos, 1);
  Py_INCREF(p0); PyTuple_SET_ITEM(t, 0, p0);
  Py_INCREF(p1); PyTuple_SET_ITEM(t, 1, p1);
  Py_DECREF(pos);
  for (int k = 0; k < 7; ++k) {
    PyObject *v = PyObject_GetAttr(env, names[k]);
    if (!v) { Py_DECREF(t); return NULL; }
    PyTuple_SET_ITEM(t, 2 + k, v);
  }
  return t;
}

static int raise_status(long status) {
  PyObject *key = PyLong_FromLong(status);
  if (!key) return -1;
  PyObject *entry = PyDict_GetItemWithError(g_errors, key);   /* borrowed */
  Py_DECREF(key);
  if (!entry) {
    if (!PyErr_Occurred()) PyErr_Format(PyExc_RuntimeError, "shipenv status %ld", status);
    return -1;
  }
  PyObject *exc = PyTuple_GetItem(entry, 0), *msg = PyTuple_GetItem(entry, 1);
  if (!exc || !msg) return -1;
  PyErr_SetObject(exc, msg);
  return -1;
}

/* The fast path of Environment._sync_world(): the same grid object, the same port tables, and the context still holds
 * this world.  1 = yes (*handle set), 0 = no / cannot tell (the Python method decides), -1 = error. */
static int world_unchanged(PyObject *env, uint64_t *handle) {
  int ok = 0;
  PyObject *slot = NULL, *sig = NULL, *a = NULL, *b = NULL, *h = NULL;
  slot = PyObject_GetAttr(env, s_slot);
  if (!slot) return -1;
  sig = PyObject_GetAttr(env, s_world_sig);
  if (!sig) goto out_err;
  if (!PyList_CheckExact(slot) || PyList_GET_SIZE(slot) != 2 || sig == Py_None || PyList_GET_ITEM(slot, 1) != sig) goto out;
  a = PyObject_GetAttr(env, s_np_game); b = PyObject_GetAttr(env, s_w_game);
  if (!a || !b) goto out_err;
  if (a != b) goto out;
  PyObject *cur[3] = {s_port_positions, s_port_fuel, s_port_cargo}, *seen[3] = {s_w_ports, s_w_fuel, s_w_cargo};
  for (int k = 0; k < 3; ++k) {
    Py_CLEAR(a); Py_CLEAR(b);
    a = PyObject_GetAttr(env, cur[k]); b = PyObject_GetAttr(env, seen[k]);
    if (!a || !b) goto out_err;
    const int eq = PyObject_RichCompareBool(a, b, Py_EQ);
    if (eq < 0) goto out_err;
    if (!eq) goto out;
  }
  h = PyObject_GetAttr(PyList_GET_ITEM(slot, 0), s_h_int);
  if (!h) { PyErr_Clear(); goto out; }
  ok = as_u64(h, handle) == 0 && *handle != 0;
out:
  Py_XDECREF(slot); Py_XDECREF(sig); Py_XDECREF(a); Py_XDECREF(b); Py_XDECREF(h);
  return ok;
out_err:
  Py_XDECREF(slot); Py_XDECREF(sig); Py_XDECREF(a); Py_XDECREF(b); Py_XDECREF(h);
  return -1;
}

/* env._sync_world() -> the context's handle (0 = could not be had: decline) */
static int context_handle(PyObject *env, uint64_t *handle) {
  const int same = world_unchanged(env, handle);
  if (same < 0) return -1;
  if (same) return 0;
  PyObject *ctx = PyObject_CallMethodNoArgs(env, s_sync_world);
  if (!ctx) return -1;                              /* its exception propagates */
  PyObject *h = PyObject_GetAttr(ctx, s_h_int);
  Py_DECREF(ctx);
  if (!h) { PyErr_Clear(); *handle = 0; return 0; }
  if (as_u64(h, handle)) *handle = 0;
  Py_DECREF(h);
  return 0;
}

static PyObject *fast_step(PyObject *self, PyObject *const *args, Py_ssize_t nargs) {
  PyObject *result = NULL;
  PyObject *ports = NULL, *dest_o = NULL, *pos = NULL, *origin_o = NULL, *cargo_o = NULL, *fuel_o = NULL, *tmp = NULL;
  if (nargs != 2) { PyErr_SetString(PyExc_TypeError, "step(env, action)"); return NULL; }
  PyObject *env = args[0], *action = args[1];
  if (!g_serve) DECLINE();

  /* ---- the action: [category, value] with plain ints ---- */
  PyObject *cat_o, *val;
  if (PyList_CheckExact(action) && PyList_GET_SIZE(action) == 2) {
    cat_o = PyList_GET_ITEM(action, 0); val = PyList_GET_ITEM(action, 1);
  } else if (PyTuple_CheckExact(action) && PyTuple_GET_SIZE(action) == 2) {
    cat_o = PyTuple_GET_ITEM(action, 0); val = PyTuple_GET_ITEM(action, 1);
  } else DECLINE();
  long cat;
  if (as_long_exact(cat_o, &cat)) DECLINE();
  ports = PyObject_GetAttr(env, s_port_positions);
  if (!ports) goto done;
  if (!PyList_CheckExact(ports) || PyList_GET_SIZE(ports) == 0) DECLINE();
  const long P = (long)PyList_GET_SIZE(ports);
  long code;
  if (cat == 1) {                                   /* MOVE_SHIP: one of the four moves of ShipMove */
    PyObject *dx_o, *dy_o;
    if (PyTuple_CheckExact(val) && PyTuple_GET_SIZE(val) == 2) {
      dx_o = PyTuple_GET_ITEM(val, 0); dy_o = PyTuple_GET_ITEM(val, 1);
    } else if (PyList_CheckExact(val) && PyList_GET_SIZE(val) == 2) {
      dx_o = PyList_GET_ITEM(val, 0); dy_o = PyList_GET_ITEM(val, 1);
    } else DECLINE();
    long dx, dy;
    if (as_long_exact(dx_o, &dx) || as_long_exact(dy_o, &dy)) DECLINE();
    if (dx == 0 && dy == -1) code = 0;
    else if (dx == -1 && dy == 0) code = 1;
    else if (dx == 0 && dy == 1) code = 2;
    else if (dx == 1 && dy == 0) code = 3;
    else DECLINE();
    dest_o = PyObject_GetAttr(env, s_destination);
    if (!dest_o) goto done;
    if (dest_o == Py_None) DECLINE();                /* "Cannot move without destination port" is Python's to raise */
  } else if (cat == 2) {                            /* SELECT_PORT */
    long v;
    if (as_long_exact(val, &v) || v < 0 || v >= P) DECLINE();
    code = 4 + v;
  } else if (cat == 3 || cat == 4) {                /* TAKE_FUEL / TAKE_CARGO */
    long v;
    if (as_long_exact(val, &v)) DECLINE();
    if (v < 0) v = 0;
    if (v > 8191) v = 8191;
    if (cat == 4) code = v < g_n_cargo ? 4 + P + v : g_ext_cargo + v;
    else code = v < g_n_fuel ? 4 + P + g_n_cargo + v : g_ext_fuel + v;
  } else DECLINE();

  /* ---- the attributes the step reads ---- */
  pos = PyObject_GetAttr(env, s_ship_position);
  if (!pos) goto done;
  long ox, oy;
  if (!PyList_CheckExact(pos) || PyList_GET_SIZE(pos) != 2 || as_long_exact(PyList_GET_ITEM(pos, 0), &ox) ||
      as_long_exact(PyList_GET_ITEM(pos, 1), &oy) || ox < 0 || ox > 255 || oy < 0 || oy > 255)
    DECLINE();
  origin_o = PyObject_GetAttr(env, s_origin);
  if (!origin_o) goto done;
  if (!dest_o) { dest_o = PyObject_GetAttr(env, s_destination); if (!dest_o) goto done; }
  long origin = NONE_PORT, dest = NONE_PORT;
  if (origin_o != Py_None && (as_long_exact(origin_o, &origin) || origin < 0 || origin >= NONE_PORT)) DECLINE();
  if (dest_o != Py_None && (as_long_exact(dest_o, &dest) || dest < 0 || dest >= NONE_PORT)) DECLINE();
  cargo_o = PyObject_GetAttr(env, s_cargo);
  if (!cargo_o) goto done;
  long cargo;
  if (as_long_exact(cargo_o, &cargo) || cargo < INT32_MIN || cargo > INT32_MAX) DECLINE();
  fuel_o = PyObject_GetAttr(env, s_fuel);
  if (!fuel_o) goto done;
  double fuel;
  int fuel_was_int = 0;
  if (PyFloat_CheckExact(fuel_o)) fuel = PyFloat_AS_DOUBLE(fuel_o);
  else if (PyLong_CheckExact(fuel_o)) {
    long f;
    if (as_long_exact(fuel_o, &f) || f < -(1L << 52) || f > (1L << 52)) DECLINE();
    fuel = (double)f; fuel_was_int = 1;
  } else DECLINE();
  uint64_t clock, word, seed, env_id, host_ptr, handle;
  PyObject *names[5] = {s_clock, s_word, s_seed, s_env_id, s_host_ptr};
  uint64_t *outs[5] = {&clock, &word, &seed, &env_id, &host_ptr};
  for (int k = 0; k < 5; ++k) {
    tmp = PyObject_GetAttr(env, names[k]);
    if (!tmp) goto done;
    if (as_u64(tmp, outs[k])) DECLINE();
    Py_CLEAR(tmp);
  }
  if (!host_ptr || clock > 0xFFFFFFFFull || word > 0xFFFFFFFFull) DECLINE();

  /* ---- the world (re-uploaded by _sync_world when it changed), then the request ---- */
  if (context_handle(env, &handle)) goto done;
  if (!handle) DECLINE();
  tmp = PyObject_GetAttr(env, s_ragged);            /* (set by _sync_world) a non-square game_size: Python's checks */
  if (!tmp) goto done;
  if (tmp != Py_False) DECLINE();
  Py_CLEAR(tmp);
  uint8_t *block = (uint8_t *)(uintptr_t)host_ptr;
  block[0] = (uint8_t)ox; block[1] = (uint8_t)oy; block[2] = (uint8_t)origin; block[3] = (uint8_t)dest;
  int32_t cargo32 = (int32_t)cargo;
  memcpy(block + 4, &cargo32, 4);
  memcpy(block + 8, &fuel, 8);
  uint32_t clock32 = (uint32_t)clock;
  memcpy(block + OFF_CLOCK, &clock32, 4);
  int16_t code16 = (int16_t)code;
  memcpy(block + OFF_ACTION, &code16, 2);
  const int rc = g_serve((void *)(uintptr_t)handle, block, seed, env_id, 0, 0u, (uint32_t)word);
  if (rc) {
    PyObject *r = PyObject_CallFunction(g_check, "i", rc);     /* raises ShipEnvError with the library's message */
    Py_XDECREF(r);
    if (!PyErr_Occurred()) PyErr_Format(PyExc_RuntimeError, "shipenv_serve_step failed (%d)", rc);
    goto done;
  }

  /* ---- the answer ---- */
  const long status = block[OFF_STATUS];
  if (status != 0) {            /* nothing changed on the device, no draw was consumed: the cell continues */
    PyObject *snap = snapshot(env);
    if (!snap) goto done;
    const int bad = PyObject_SetAttr(env, s_spec, snap);
    Py_DECREF(snap);
    if (bad) goto done;
    raise_status(status);
    goto done;
  }
  const long nx = block[0], ny = block[1], norigin = block[2], ndest = block[3];
  int32_t ncargo;
  double nfuel, reward;
  uint32_t nclock, done_word;
  memcpy(&ncargo, block + 4, 4);
  memcpy(&nfuel, block + 8, 8);
  memcpy(&reward, block + OFF_REWARD, 8);
  memcpy(&nclock, block + OFF_CLOCK, 4);
  memcpy(&done_word, block + OFF_DONE, 4);
  const int moved = nx != ox || ny != oy;
  const int keep_int = fuel_was_int && !(cat == 1 && moved);

#define SET_ATTR(name, expr) do { PyObject *v_ = (expr); if (!v_) goto done; \
    const int bad_ = PyObject_SetAttr(env, name, v_); Py_DECREF(v_); if (bad_) goto done; } while (0)
  SET_ATTR(s_clock, PyLong_FromUnsignedLong(nclock));
  SET_ATTR(s_word, PyLong_FromLong(0));
  if (norigin == NONE_PORT) { if (PyObject_SetAttr(env, s_origin, Py_None)) goto done; }
  else SET_ATTR(s_origin, PyLong_FromLong(norigin));
  if (ndest == NONE_PORT) { if (PyObject_SetAttr(env, s_destination, Py_None)) goto done; }
  else SET_ATTR(s_destination, PyLong_FromLong(ndest));
  SET_ATTR(s_cargo, PyLong_FromLong(ncargo));
  if (keep_int && nfuel == (double)(long long)nfuel) SET_ATTR(s_fuel, PyLong_FromLongLong((long long)nfuel));
  else SET_ATTR(s_fuel, PyFloat_FromDouble(nfuel));
  if (moved) {
    SET_ATTR(s_ship_position, Py_BuildValue("[ll]", nx, ny));                 /* a fresh list, like :297 */
    PyObject *grid = PyObject_GetAttr(env, s_np_game);                           /* :303-304 */
    if (!grid) goto done;
    PyObject *k_old = Py_BuildValue("(ll)", ox, oy), *k_new = Py_BuildValue("(ll)", nx, ny);
    PyObject *v_old = PyLong_FromLong(ENT_TRAVEL), *v_new = PyLong_FromLong(ENT_BOAT);
    int bad = !k_old || !k_new || !v_old || !v_new;
    if (!bad) bad = PyObject_SetItem(grid, k_old, v_old) || PyObject_SetItem(grid, k_new, v_new);
    Py_XDECREF(k_old); Py_XDECREF(k_new); Py_XDECREF(v_old); Py_XDECREF(v_new); Py_DECREF(grid);
    if (bad) goto done;
  }
  {
    PyObject *snap = snapshot(env);
    if (!snap) goto done;
    const int bad = PyObject_SetAttr(env, s_spec, snap);
    Py_DECREF(snap);
    if (bad) goto done;
  }
  {
    PyObject *state = PyObject_CallMethodNoArgs(env, s_build_state);
    if (!state) goto done;
    PyObject *meta = PyDict_New(), *rew = PyFloat_FromDouble(reward);
    if (!meta || !rew) { Py_DECREF(state); Py_XDECREF(meta); Py_XDECREF(rew); goto done; }
    result = PyTuple_Pack(4, state, rew, (done_word & 1u) ? Py_True : Py_False, meta);
    Py_DECREF(state); Py_DECREF(meta); Py_DECREF(rew);
  }
done:
  Py_XDECREF(ports); Py_XDECREF(dest_o); Py_XDECREF(pos); Py_XDECREF(origin_o); Py_XDECREF(cargo_o);
  Py_XDECREF(fuel_o); Py_XDECREF(tmp);
  return result;
}

/* sample_action() when the previous step() already drew it for exactly this state (shipenv_step_sample's look-ahead) */
static PyObject *fast_sample_action(PyObject *self, PyObject *env) {
  PyObject *result = NULL, *spec = NULL, *snap = NULL, *ports = NULL, *tmp = NULL;
  uint64_t handle, host_ptr, word;
  if (!g_serve) DECLINE();
  if (context_handle(env, &handle)) goto done;
  spec = PyObject_GetAttr(env, s_spec);
  if (!spec) goto done;
  if (spec == Py_None) DECLINE();
  snap = snapshot(env);
  if (!snap) goto done;
  const int eq = PyObject_RichCompareBool(spec, snap, Py_EQ);
  if (eq < 0) goto done;
  if (!eq) DECLINE();
  tmp = PyObject_GetAttr(env, s_host_ptr);
  if (!tmp) goto done;
  if (as_u64(tmp, &host_ptr) || !host_ptr) DECLINE();
  Py_CLEAR(tmp);
  tmp = PyObject_GetAttr(env, s_word);
  if (!tmp) goto done;
  if (as_u64(tmp, &word)) DECLINE();
  Py_CLEAR(tmp);
  ports = PyObject_GetAttr(env, s_port_positions);
  if (!ports) goto done;
  if (!PyList_CheckExact(ports)) DECLINE();
  const long P = (long)PyList_GET_SIZE(ports);
  if (PyObject_SetAttr(env, s_spec, Py_None)) goto done;       /* the draw is used once */
  const uint8_t *block = (const uint8_t *)(uintptr_t)host_ptr;
  int16_t taken;
  memcpy(&taken, block + OFF_NEXT, 2);
  const long words = block[OFF_NEXT + 2], status = block[OFF_NEXT + 3];
  if (status != 0) { raise_status(status); goto done; }
  {
    PyObject *w = PyLong_FromUnsignedLongLong(word + (uint64_t)words);
    if (!w) goto done;
    const int bad = PyObject_SetAttr(env, s_word, w);
    Py_DECREF(w);
    if (bad) goto done;
  }
  const long a = taken;
  if (a < 0) { PyErr_Format(PyExc_RuntimeError, "shipenv: look-ahead action %ld", a); goto done; }
  if (a < 4) result = Py_BuildValue("[lO]", 1L, PyTuple_GET_ITEM(g_moves, a));            /* MOVE_SHIP */
  else if (a < 4 + P) result = Py_BuildValue("[ll]", 2L, a - 4);                          /* SELECT_PORT */
  else if (a < 4 + P + g_n_cargo) result = Py_BuildValue("[ll]", 4L, a - (4 + P));        /* TAKE_CARGO */
  else if (a < 4 + P + g_n_cargo + g_n_fuel) result = Py_BuildValue("[ll]", 3L, a - (4 + P + g_n_cargo));   /* TAKE_FUEL */
  else if (a < g_ext_fuel) result = Py_BuildValue("[ll]", 4L, a - g_ext_cargo);
  else result = Py_BuildValue("[ll]", 3L, a - g_ext_fuel);
done:
  Py_XDECREF(spec); Py_XDECREF(snap); Py_XDECREF(ports); Py_XDECREF(tmp);
  return result;
}

/* bind(serve_step address, errors dict, check callable, MOVES tuple, n_cargo, n_fuel, ext_cargo, ext_fuel) */
static PyObject *fast_bind(PyObject *self, PyObject *args) {
  unsigned long long addr;
  PyObject *errors, *check, *moves;
  long n_cargo, n_fuel, ext_cargo, ext_fuel;
  if (!PyArg_ParseTuple(args, "KO!OO!llll", &addr, &PyDict_Type, &errors, &check, &PyTuple_Type, &moves, &n_cargo,
                        &n_fuel, &ext_cargo, &ext_fuel))
    return NULL;
  if (PyTuple_GET_SIZE(moves) != 4 || !PyCallable_Check(check)) {
    PyErr_SetString(PyExc_TypeError, "bind: four moves and a callable check are needed");
    return NULL;
  }
  Py_INCREF(errors); Py_XSETREF(g_errors, errors);
  Py_INCREF(check); Py_XSETREF(g_check, check);
  Py_INCREF(moves); Py_XSETREF(g_moves, moves);
  g_n_cargo = n_cargo; g_n_fuel = n_fuel; g_ext_cargo = ext_cargo; g_ext_fuel = ext_fuel;
  g_serve = (serve_fn_t)(uintptr_t)addr;
  Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"step", (PyCFunction)(void (*)(void))fast_step, METH_FASTCALL, "step(env, action) or NotImplemented"},
    {"sample_action", fast_sample_action, METH_O, "sample_action(env) or NotImplemented"},
    {"bind", fast_bind, METH_VARARGS, "bind(serve_step address, errors, check, moves, n_cargo, n_fuel, ext_cargo, ext_fuel)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_facade_fast",
                                    "C fast path of shippingenv_b200.Environment.step / sample_action", -1, methods};

PyMODINIT_FUNC PyInit__facade_fast(void) {
#define INTERN(var, text) do { var = PyUnicode_InternFromString(text); if (!var) return NULL; } while (0)
  INTERN(s_port_positions, "port_positions"); INTERN(s_destination, "destination_port_index");
  INTERN(s_origin, "origin_port_index"); INTERN(s_ship_position, "ship_position"); INTERN(s_cargo, "cargo");
  INTERN(s_fuel, "fuel"); INTERN(s_clock, "_clock"); INTERN(s_word, "_word"); INTERN(s_seed, "_seed");
  INTERN(s_env_id, "_env_id"); INTERN(s_host_ptr, "_host_ptr"); INTERN(s_ragged, "_ragged");
  INTERN(s_sync_world, "_sync_world"); INTERN(s_h_int, "_h_int"); INTERN(s_spec, "_spec");
  INTERN(s_world_sig, "_world_sig"); INTERN(s_np_game, "np_game"); INTERN(s_build_state, "_build_state");
  INTERN(s_slot, "_slot"); INTERN(s_w_game, "_w_game"); INTERN(s_w_ports, "_w_ports"); INTERN(s_w_fuel, "_w_fuel");
  INTERN(s_w_cargo, "_w_cargo"); INTERN(s_port_fuel, "port_fuel"); INTERN(s_port_cargo, "port_cargo");
  return PyModule_Create(&module);
}
