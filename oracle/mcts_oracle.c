/* oracle/mcts_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's tree search for ONE search thread:
 *   MCTSAI_T::act / advanceMoves          src_cpp/elf/ai/tree_search/mcts.h:59-81,155-167
 *   TreeSearchT::run / chooseAction       src_cpp/elf/ai/tree_search/tree_search.h:410-426,495-528
 *   batch_rollouts / single_rollout       tree_search.h:201-322
 *   NodeT::findMove/UCT/addVirtualLoss/followEdge/setEvaluation/updateEdgeStats
 *                                         tree_search_node.h:176-302,361-397
 *   EdgeInfo::getScore                    tree_search_base.h:132-157
 *   SearchTreeT::treeAdvance              tree_search_node.h:420-436
 *   MCTSActor::pre_evaluate / remove_pass_if_dangerous / pi2response / normalize
 *                                         src_cpp/elfgames/go/mcts/mcts.h:185-332
 * on top of the board restatement (go_oracle.c).  Pinned against the compiled reference search
 * (oracle/_ref, ref_mcts_shim.cc) in tests/test_mcts_oracle_vs_ref.py.
 *
 * What the reference leaves to its containers:
 *  - edges are STORED in the order pi2response produces them (descending prior), but every arg-max
 *    the reference takes by iterating its unordered_map with a strict '>' -- the PUCT choice of each
 *    descent step (NodeT::UCT) and the most-visited choice at the root -- is the FIRST maximum in
 *    that container's iteration order: exact score ties (frequent once priors fall below the
 *    rounding of q, or without the prior term) are resolved through container_order below;
 *  - unique leaves of a batch are backed up in first-occurrence order (the reference iterates an
 *    unordered_map keyed by node address): only the last bits of reward sums can differ.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fakenet.h"
#include "go_oracle.h"
#include "stdsort_emul.h"

typedef void (*mo_eval_cb)(int n, const float* feats, const uint64_t* hashes, float* pi, float* v);

typedef struct {
  float prior, reward, vloss;
  int visits, child, action;
} Edge;

typedef struct {
  GoOracle* state; /* NULL until allocateState */
  int status;      /* 0 not visited, 1 eval requested, 2 visited */
  int num_visits;
  float V, mean_q, parent_q;
  int flip;
  Edge* edges;
  int n_edges;
  int n_touched; /* edges [0, n_touched) have been selected at least once (prefix property check) */
  int fullscan;  /* a tie was resolved in favour of an edge beyond the prefix: the CUDA path scans all edges from then on */
  int parent, parent_edge;
  int alive;
} Node;

typedef struct MctsOracle {
  int N;
  int R, B, vl, persistent, use_prior, uqz, ruqz, ply_pass_enabled, remove_pass;
  int std_sort_ties; /* equal probabilities in std::sort's order (pi2response) instead of by ascending move */
  float c_puct, komi;
  mo_eval_cb cb;
  Node* nodes;
  int n_nodes, cap_nodes;
  int root;
  int next_move_number;
  long n_evals;
  long prefix_violations; /* times the arg-max over [0, n_touched] differed from the full arg-max */
  long prefix_checks;
  long tie_breaks;     /* descent steps whose maximum was tied and went through the container order */
  long tie_breaks_out; /* ... of which the chosen edge lay beyond the first never-selected one */
} MctsOracle;

static int add_node(MctsOracle* m, float parent_q, int parent, int parent_edge) {
  if (m->n_nodes == m->cap_nodes) {
    m->cap_nodes = m->cap_nodes ? 2 * m->cap_nodes : 1024;
    m->nodes = (Node*)realloc(m->nodes, sizeof(Node) * (size_t)m->cap_nodes);
  }
  Node* nd = &m->nodes[m->n_nodes];
  memset(nd, 0, sizeof(*nd));
  nd->parent_q = parent_q;
  nd->mean_q = parent_q; /* NodeT ctor, tree_search_node.h:98-103 */
  nd->parent = parent;
  nd->parent_edge = parent_edge;
  nd->alive = 1;
  return m->n_nodes++;
}

static void free_node(MctsOracle* m, int id) {
  Node* nd = &m->nodes[id];
  if (!nd->alive) return;
  if (nd->state) go_free(nd->state);
  free(nd->edges);
  nd->state = NULL;
  nd->edges = NULL;
  nd->n_edges = 0;
  nd->alive = 0;
}

static void recursive_free(MctsOracle* m, int id) { /* tree_search_node.h:457-467 */
  if (id < 0) return;
  Node* nd = &m->nodes[id];
  for (int i = 0; i < nd->n_edges; ++i) recursive_free(m, nd->edges[i].child);
  free_node(m, id);
}

static void tree_clear(MctsOracle* m) { /* SearchTreeT::clear, tree_search_node.h:413-418 */
  for (int i = 0; i < m->n_nodes; ++i) free_node(m, i);
  m->n_nodes = 0;
  m->root = add_node(m, 0.0f, -1, 0);
  m->next_move_number = 0;
}

static void tree_advance(MctsOracle* m, int action) { /* tree_search_node.h:420-436 */
  Node* r = &m->nodes[m->root];
  int next_root = -1;
  for (int i = 0; i < r->n_edges; ++i) {
    if (r->edges[i].action == action)
      next_root = r->edges[i].child;
    else
      recursive_free(m, r->edges[i].child);
  }
  free_node(m, m->root);
  if (next_root < 0) next_root = add_node(m, 0.0f, -1, 0);
  m->root = next_root;
  m->nodes[m->root].parent = -1;
}

MctsOracle* mo_new(int N, const int32_t* iopts, const float* fopts, mo_eval_cb cb) {
  MctsOracle* m = (MctsOracle*)calloc(1, sizeof(MctsOracle));
  m->N = N;
  m->R = iopts[0];
  m->B = iopts[1];
  m->vl = iopts[2];
  m->persistent = iopts[3];
  m->use_prior = iopts[4];
  m->uqz = iopts[5];
  m->ruqz = iopts[6];
  m->ply_pass_enabled = iopts[7];
  m->remove_pass = iopts[8];
  m->std_sort_ties = iopts[11]; /* [9] seed, [10] threads: unused here */
  m->c_puct = fopts[0];
  m->komi = fopts[1];
  m->cb = cb;
  tree_clear(m);
  return m;
}

void mo_free(MctsOracle* m) {
  if (!m) return;
  for (int i = 0; i < m->n_nodes; ++i) free_node(m, i);
  free(m->nodes);
  free(m);
}

long mo_num_evals(const MctsOracle* m) { return m->n_evals; }

/* The CUDA search only scans the selected prefix of the prior-sorted edges plus the first TWO
 * never-selected ones (k_select) and falls back to a full scan with the container-order tie-break
 * when the maximum over that range is tied; this restatement scans everything, like the reference,
 * and counts how often the short scan's untied maximum would not be the full answer (must be 0). */
long mo_prefix_violations(const MctsOracle* m) { return m->prefix_violations; }
long mo_prefix_checks(const MctsOracle* m) { return m->prefix_checks; }
long mo_tie_breaks(const MctsOracle* m) { return m->tie_breaks; }
long mo_tie_breaks_beyond_prefix(const MctsOracle* m) { return m->tie_breaks_out; }

/* ---- evaluation: MCTSActor::evaluate for one state (go/mcts/mcts.h:73-121,185-332) ---- */
typedef struct {
  float p;
  int a;
} Cand;

static int cand_cmp(const void* x, const void* y) {
  const Cand* a = (const Cand*)x;
  const Cand* b = (const Cand*)y;
  if (a->p > b->p) return -1;
  if (a->p < b->p) return 1;
  return a->a - b->a;
}

static void evaluate_node(MctsOracle* m, Node* nd) {
  const GoOracle* s = nd->state;
  const int N = m->N, P = N * N;
  nd->flip = go_next_player(s) == 2; /* q_flip, mcts.h:186 */
  nd->n_edges = 0;
  if (go_terminated(s)) { /* pre_evaluate, mcts.h:188-203 */
    float fv = go_evaluate(s, m->komi);
    nd->V = fv > 0 ? 1.0f : -1.0f;
    nd->status = 2;
    return;
  }
  float* pi = (float*)malloc(sizeof(float) * (size_t)(P + 1));
  float v;
  uint64_t h = go_hash(s);
  if (m->cb) {
    float* feats = (float*)malloc(sizeof(float) * 18 * (size_t)P);
    go_features_agz(s, 0, feats);
    m->cb(1, feats, &h, pi, &v);
    free(feats);
  } else {
    for (int a = 0; a <= P; ++a) pi[a] = fakenet_pi(h, a);
    v = fakenet_value(h);
  }
  m->n_evals++;
  /* post_nn_result, mcts.h:209-230 */
  int32_t info[12];
  go_info(s, info);
  int pass_enabled = info[0] >= m->ply_pass_enabled;
  if (m->remove_pass && pass_enabled && info[4] != P) { /* remove_pass_if_dangerous, mcts.h:232-242 */
    int black_win = go_evaluate(s, m->komi) > 0;
    if ((black_win && info[1] == 2) || (!black_win && info[1] == 1)) pass_enabled = 0;
  }
  /* pi2response, mcts.h:256-332 (rotation_flip off: action2Coord is the identity) */
  Cand* c = (Cand*)malloc(sizeof(Cand) * (size_t)(P + 1));
  int nc = 0;
  for (int a = 0; a <= P; ++a) {
    int valid = (a == P) ? pass_enabled : go_check_move(s, a);
    if (valid) {
      c[nc].p = pi[a];
      c[nc].a = a;
      nc++;
    }
  }
  qsort(c, (size_t)nc, sizeof(Cand), cand_cmp);
  int tie = 0;
  for (int i = 1; i < nc; ++i) tie |= c[i].p == c[i - 1].p;
  if (m->std_sort_ties && tie) {
    /* pi2response as written: ALL P+1 pairs in network-action order through std::sort (comparator on the
     * probability only), then the validity filter -- equal probabilities end up in libstdc++'s order */
    uint32_t keys[512];
    uint16_t ord[512];
    for (int a = 0; a <= P; ++a) {
      memcpy(&keys[a], &pi[a], 4);
      ord[a] = (uint16_t)a;
    }
    SseCtx sc = {keys, 0};
    sse_sort(&sc, ord, P + 1);
    nc = 0;
    for (int i = 0; i <= P; ++i) {
      const int a = ord[i];
      if ((a == P) ? pass_enabled : go_check_move(s, a)) {
        c[nc].p = pi[a];
        c[nc].a = a;
        nc++;
      }
    }
  }
  if (nc == 0 && !pass_enabled) {
    c[0].p = 1.0f;
    c[0].a = P;
    nc = 1;
  }
  float total = 1e-10f; /* normalize, mcts.h:244-254 */
  for (int i = 0; i < nc; ++i) total += c[i].p;
  nd->edges = (Edge*)calloc((size_t)nc, sizeof(Edge));
  for (int i = 0; i < nc; ++i) {
    nd->edges[i].prior = c[i].p / total;
    nd->edges[i].action = c[i].a;
    nd->edges[i].child = -1;
  }
  nd->n_edges = nc;
  nd->V = v;
  nd->status = 2; /* setEvaluation, tree_search_node.h:176-203 */
  free(c);
  free(pi);
}

static void container_order(int N, const Edge* edges, int n, int* order);

/* ---- one wave: batch_rollouts, tree_search.h:201-262 ---- */
static void batch_rollouts(MctsOracle* m) {
  const int B = m->B;
  int* leaves = (int*)malloc(sizeof(int) * (size_t)B);
  for (int j = 0; j < B; ++j) {
    int node = m->root, depth = 0;
    while (m->nodes[node].status == 2) { /* single_rollout, tree_search.h:265-322 */
      Node* nd = &m->nodes[node];
      if (nd->n_edges == 0) break; /* findMove returns false */
      if (m->uqz || (m->ruqz && depth == 0)) nd->mean_q = 0.0f;
      /* UCT, tree_search_node.h:361-397 + getScore, tree_search_base.h:132-157 */
      const double sq = sqrt((double)(nd->num_visits + 1));
      float best = -FLT_MAX, tuq = 0.0f, best_prefix = -FLT_MAX;
      int besti = -1, tv = 0, besti_prefix = -1, ties = 0, ties_prefix = 0;
      const int lim = (nd->fullscan || nd->n_touched + 2 >= nd->n_edges) ? nd->n_edges : nd->n_touched + 2;
      float scores[512];
      for (int i = 0; i < nd->n_edges; ++i) {
        const Edge* e = &nd->edges[i];
        float r = nd->flip ? -e->reward : e->reward;
        r -= e->vloss;
        const int nwl = (int)((float)e->visits + e->vloss);
        const float q = nwl > 0 ? r / (float)nwl : (nd->flip ? -nd->mean_q : nd->mean_q);
        const float uq = e->visits > 0 ? e->reward / (float)e->visits : nd->mean_q;
        const float u = (float)((double)(e->prior / (float)(1 + e->visits)) * sq);
        const float score = m->use_prior ? fmaf(u, m->c_puct, q) : q;
        scores[i] = score;
        if (score > best) {
          best = score;
          besti = i;
          ties = 1;
        } else if (score == best) {
          ties++;
        }
        if (i < lim) {
          if (score > best_prefix) {
            best_prefix = score;
            besti_prefix = i;
            ties_prefix = 1;
          } else if (score == best_prefix) {
            ties_prefix++;
          }
        }
        if (nwl != 0 && i >= nd->n_touched && !nd->fullscan) m->prefix_violations++; /* a touched edge outside the prefix */
        if (nwl != 0) {
          tuq += uq;
          tv++;
        }
      }
      if (ties > 1) { /* NodeT::UCT walks the unordered_map with a strict '>': first maximum in ITS order */
        int order[512];
        m->tie_breaks++;
        container_order(m->N, nd->edges, nd->n_edges, order);
        for (int k = 0; k < nd->n_edges; ++k)
          if (scores[order[k]] == best) {
            besti = order[k];
            break;
          }
      }
      m->prefix_checks++;
      /* the short scan is trusted only when its maximum is untied: then it must be the answer */
      if (ties_prefix == 1 && (besti_prefix != besti || ties != 1)) m->prefix_violations++;
      if (besti > nd->n_touched && !nd->fullscan) {
        nd->fullscan = 1;
        m->tie_breaks_out++;
      }
      if (besti == nd->n_touched) nd->n_touched++;
      nd->mean_q = (nd->parent_q + tuq) / (float)(tv + 1); /* findMove, tree_search_node.h:227 */
      Edge* e = &nd->edges[besti];
      if (m->vl > 0) e->vloss += (float)m->vl; /* addVirtualLoss */
      if (e->child < 0) {                      /* followEdge: addNode(unsignedMeanQ_) */
        int id = add_node(m, nd->mean_q, node, besti);
        nd = &m->nodes[node]; /* realloc may have moved the pool */
        e = &nd->edges[besti];
        e->child = id;
      }
      Node* ch = &m->nodes[e->child];
      if (!ch->state) { /* allocateState, tree_search.h:175-190 */
        ch->state = go_clone(nd->state);
        if (!go_forward(ch->state, e->action)) {
          go_free(ch->state);
          ch->state = NULL;
          break;
        }
      }
      node = e->child;
      depth++;
    }
    leaves[j] = node;
  }
  /* claim + evaluate (requestEvaluation / actor.evaluate / setEvaluation) */
  for (int j = 0; j < B; ++j) {
    Node* nd = &m->nodes[leaves[j]];
    if (nd->status == 0) {
      nd->status = 1;
      evaluate_node(m, nd);
    }
  }
  /* backup once per unique leaf, first-occurrence order (tree_search.h:245-259) */
  for (int j = 0; j < B; ++j) {
    int count = 0, first = 1;
    for (int k = 0; k < B; ++k)
      if (leaves[k] == leaves[j]) {
        if (k < j) first = 0;
        count++;
      }
    if (!first) continue;
    const float reward = m->nodes[leaves[j]].V;
    int node = leaves[j];
    while (m->nodes[node].parent >= 0) { /* updateEdgeStats, tree_search_node.h:253-278 */
      const int p = m->nodes[node].parent;
      Edge* e = &m->nodes[p].edges[m->nodes[node].parent_edge];
      m->nodes[p].num_visits++;
      e->reward += reward;
      e->visits++;
      e->vloss -= (float)m->vl * (float)count;
      node = p;
    }
  }
  free(leaves);
}

/* Iteration order of the reference's edge container, std::unordered_map<Coord, EdgeInfo>
 * (tree_search_node.h:310), after NodeT::setEvaluation (:176-203) inserted the edges in storage order.
 * libstdc++ keeps one forward list: a key whose bucket is empty goes to the front of the list, a key
 * whose bucket is in use goes right behind the node before that bucket's first node; hash = identity,
 * bucket = key % count, counts 13, 29, 59, 127, 257, 541 growing when key 14, 30, 60, 128, 258
 * arrives; a rehash re-inserts the list front to back by the same rules.  order[i] = storage index of
 * the i-th edge visited.  Pinned against the compiled reference (ref_mcts_last_order) in
 * tests/test_mcts_oracle_vs_ref.py. */
#define CO_NIL (-1)
#define CO_HEAD (-2)
#define CO_EMPTY (-3)
static void co_link(int* nxt, int* bkt, int* head, int nb, int key) {
  const int b = key % nb, before = bkt[b];
  if (before == CO_EMPTY) {
    nxt[key] = *head;
    if (*head != CO_NIL) bkt[*head % nb] = key;
    *head = key;
    bkt[b] = CO_HEAD;
  } else if (before == CO_HEAD) {
    nxt[key] = *head;
    *head = key;
  } else {
    nxt[key] = nxt[before];
    nxt[before] = key;
  }
}
static void container_order(int N, const Edge* edges, int n, int* order) {
  int nxt[448], idx[448], bkt[544];
  int nb = 13, head = CO_NIL;
  for (int b = 0; b < nb; ++b) bkt[b] = CO_EMPTY;
  for (int i = 0; i < n; ++i) {
    if (i == 13 || i == 29 || i == 59 || i == 127 || i == 257) {
      nb = i == 13 ? 29 : i == 29 ? 59 : i == 59 ? 127 : i == 127 ? 257 : 541;
      for (int b = 0; b < nb; ++b) bkt[b] = CO_EMPTY;
      int p = head;
      head = CO_NIL;
      while (p != CO_NIL) {
        const int q = nxt[p];
        co_link(nxt, bkt, &head, nb, p);
        p = q;
      }
    }
    const int a = edges[i].action;
    const int key = a >= N * N ? 0 : ((a % N) + 1) * (N + 2) + (a / N) + 1; /* board.h:183-184, M_PASS = 0 */
    idx[key] = i;
    co_link(nxt, bkt, &head, nb, key);
  }
  int k = 0;
  for (int p = head; p != CO_NIL; p = nxt[p]) order[k++] = idx[p];
}

/* the root edges' actions in container order (as ref_mcts_last_order); returns their number */
int mo_last_order(MctsOracle* m, int32_t* actions) {
  const Node* root = &m->nodes[m->root];
  int order[512];
  container_order(m->N, root->edges, root->n_edges, order);
  for (int i = 0; i < root->n_edges; ++i) actions[i] = root->edges[order[i]].action;
  return root->n_edges;
}

/* MCTSAI_T::act, mcts.h:59-81.  Outputs by action index as ref_mcts_act. */
int mo_act(MctsOracle* m, const GoOracle* s, int32_t* visits, float* wsum, float* prior,
           float* root_value, float* best_q, int32_t* total_visits) {
  const int P1 = m->N * m->N + 1;
  /* align_state / advanceMoves, mcts.h:139-167 */
  if (!m->persistent) {
    tree_clear(m);
  } else {
    int nm = go_num_moves(s);
    if (m->next_move_number > nm) {
      tree_clear(m);
    } else {
      for (int i = m->next_move_number; i < nm; ++i) tree_advance(m, go_move_at(s, i));
      m->next_move_number = nm;
    }
  }
  Node* root = &m->nodes[m->root];
  if (!root->state) root->state = go_clone(s); /* setRootNodeState */
  for (int idx = 0; idx < m->R; idx += m->B) batch_rollouts(m);
  root = &m->nodes[m->root];
  for (int a = 0; a < P1; ++a) {
    if (visits) visits[a] = -1;
    if (wsum) wsum[a] = 0;
    if (prior) prior[a] = 0;
  }
  int best = -1, bestn = -1, tot = 0;
  int order[512];
  container_order(m->N, root->edges, root->n_edges, order);
  for (int k = 0; k < root->n_edges; ++k) { /* addActions, tree_search_base.h:237-294: container order */
    const int i = order[k];
    const Edge* e = &root->edges[i];
    if (visits) visits[e->action] = e->visits;
    if (wsum) wsum[e->action] = e->reward;
    if (prior) prior[e->action] = e->prior;
    tot += e->visits;
    if (e->visits > bestn) {
      bestn = e->visits;
      best = i;
    }
  }
  if (root_value) *root_value = root->V;
  if (total_visits) *total_visits = tot;
  if (best_q) *best_q = (tot == 0 || best < 0) ? root->V : root->edges[best].reward / (float)root->edges[best].visits;
  return best >= 0 ? root->edges[best].action : -1;
}
