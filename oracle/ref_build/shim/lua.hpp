#pragma once
struct lua_State;
