"""B200-native BSMS processor (hot path of Eydcao/BSMS-GNN)."""
