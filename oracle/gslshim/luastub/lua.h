/* Lua stub — TEST INFRASTRUCTURE ONLY.  readConf.h (pulled in by examples/target/sampleTarget.h) names
 * lua_State; Lua is absent from this image and nothing on the PMC hot path calls it. */
#ifndef PMCB200_LUA_STUB_H
#define PMCB200_LUA_STUB_H
typedef struct lua_State lua_State;
const char *lua_tostring(lua_State *L, int idx);
#endif
