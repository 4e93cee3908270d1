// placeholder: MCTS reference shim (filled in with the search path)
