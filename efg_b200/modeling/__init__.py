"""Model building blocks of the hot path (mirror of the efg.modeling names the playground imports)."""
