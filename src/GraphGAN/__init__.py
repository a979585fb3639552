"""Flat drop-in aliases of graphgan_b200.{config,generator,discriminator,graph_gan} under the reference's directory
layout: `import config`, `import generator`, ... from this directory resolve to the B200 build."""
