// Fan-out of C-ABI calls over the child contexts of a multi-GPU ctx (snpg_create_multi; exchange.cu).
#pragma once

#include "snpg_ctx.h"

namespace snpg {

inline bool is_team(const snpg_ctx *ctx) { return ctx && !ctx->team.empty(); }
int team_fail(snpg_ctx *team, snpg_ctx *child, int rc);      // the child's error text becomes the team's
int team_wire(snpg_ctx *team, int64_t max_bootstraps);       // (re)builds the children's exchange regions and connects them

}  // namespace snpg

#define TEAM_UNSUPPORTED(ctx, name) \
    do { if (snpg::is_team(ctx)) return snpg::fail(ctx, SNPG_E_STATE, name " is not available on a multi-GPU ctx (snpg_create_multi)"); } while (0)
// runs `call` (an expression using the child `c`) on the first GPU / on every GPU and returns
#define TEAM_FORWARD0(ctx, call) \
    do { if (snpg::is_team(ctx)) { snpg_ctx *c = (ctx)->team[0]; const int rc__ = (call); return rc__ ? snpg::team_fail(ctx, c, rc__) : rc__; } } while (0)
#define TEAM_EACH(ctx, call) \
    do { if (snpg::is_team(ctx)) { for (snpg_ctx *c : (ctx)->team) { const int rc__ = (call); if (rc__) return snpg::team_fail(ctx, c, rc__); } return SNPG_OK; } } while (0)
