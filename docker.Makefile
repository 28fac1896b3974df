# Images for internevo_b200 (B200 / sm_100a; needs CUDA >= 12.8 for the compute_100a target).
#   make -f docker.Makefile devel-ubuntu     full image: toolchain + built extension + test deps
#   make -f docker.Makefile runtime-ubuntu   slim image: python package + _C.so only
#   make -f docker.Makefile devel-rocky      same on a RHEL-family base
#   make -f docker.Makefile experiment       next-toolchain test image (experiment/README.md)
DOCKER_REGISTRY ?= docker.io
DOCKER_ORG      ?= internevo-b200
DOCKER_IMAGE    ?= internevo_b200
CUDA_VERSION    ?= 12.9
TORCH_IMAGE     ?= nvcr.io/nvidia/pytorch:25.03-py3
VERSION         := $(shell cat version.txt)
TAG             ?= $(VERSION)-cuda$(CUDA_VERSION)
BUILD           ?= docker build --progress=plain
NAME             = $(DOCKER_REGISTRY)/$(DOCKER_ORG)/$(DOCKER_IMAGE)

.PHONY: all devel-ubuntu runtime-ubuntu devel-rocky experiment push clean
all: devel-ubuntu

devel-ubuntu:
	$(BUILD) --build-arg BASE=$(TORCH_IMAGE) --target devel -t $(NAME):$(TAG)-devel-ubuntu -f docker/Dockerfile .

runtime-ubuntu:
	$(BUILD) --build-arg BASE=$(TORCH_IMAGE) --target runtime -t $(NAME):$(TAG)-runtime-ubuntu -f docker/Dockerfile .

devel-rocky:
	$(BUILD) --build-arg CUDA_VERSION=$(CUDA_VERSION) -t $(NAME):$(TAG)-devel-rocky -f docker/Dockerfile-rocky .

experiment:
	$(BUILD) --build-arg BASE=$(TORCH_IMAGE) -t $(NAME):$(VERSION)-experiment -f experiment/Dockerfile-next .

push:
	docker push $(NAME):$(TAG)-devel-ubuntu
	-docker push $(NAME):$(TAG)-runtime-ubuntu

clean:
	-docker rmi $(NAME):$(TAG)-devel-ubuntu $(NAME):$(TAG)-runtime-ubuntu $(NAME):$(TAG)-devel-rocky
