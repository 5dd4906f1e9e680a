"""Host-side mirror of the reference's `models` package for the render hot path only
(rendering__, star__, nerf, resnet, embedder, types__)."""
